// No C++ exception may cross the C boundary: every extern "C" entry point runs its body inside guarded().
#pragma once
#include <exception>
#include <new>
#include <string>
#include "../../include/adypt_b200.h"

namespace adypt {

int fail(int code, const std::string &msg);

template <class F> inline int guarded(F &&body) noexcept
{
	try {
		return body();
	} catch (const std::bad_alloc &) {
		try { return fail(ADYPT_ENOMEM, "out of host memory"); } catch (...) { return ADYPT_ENOMEM; }
	} catch (const std::exception &e) {
		try { return fail(ADYPT_EINTERNAL, std::string("internal error: ") + e.what()); } catch (...) { return ADYPT_EINTERNAL; }
	} catch (...) {
		return ADYPT_EINTERNAL;
	}
}

} // namespace adypt
