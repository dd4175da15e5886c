// Test harness (not part of the product): decodes every file of a directory with the library's texture decoders.
// tests/test_textures.py builds it with -fsanitize=address,undefined and feeds it mutated image files.
#include <cstdio>
#include <new>
#include <vector>
#include <string>
#include <dirent.h>
#include "host/host_scene.h"
int main(int argc, char **argv)
{
	DIR *d = opendir(argv[1]);
	int ok = 0, bad = 0;
	while (dirent *e = readdir(d)) {
		if (e->d_name[0] == '.') continue;
		std::string p = std::string(argv[1]) + "/" + e->d_name;
		adypt::host::DecodedImage img;
		bool decoded = false;
		try { decoded = adypt::host::decode_image_file(p.c_str(), &img); }
		catch (const std::bad_alloc &) { decoded = false; } // a header that asks for gigabytes: the C-ABI reports ADYPT_ENOMEM
		if (decoded) { ++ok; if ((size_t)img.width * img.height * 3 != img.rgb.size()) { printf("SIZE MISMATCH %s\n", p.c_str()); return 1; } }
		else ++bad;
	}
	printf("decoded %d refused %d\n", ok, bad);
	return 0;
}
