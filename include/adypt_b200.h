/*
 * adypt_b200 -- C-ABI of the B200-native ray-traversal / path-tracing core for AdamYuan/Adypt.
 *
 * This library replaces the reference's OpenGL 4.5 compute path (src/Tracer/OglScene.*,
 * src/Tracer/OglPathTracer.*, shaders/{traversal,primaryray,pathtracer}.glsl) and nothing else.
 * Adypt has no FFI of its own; the seam is the C++ class surface Instance uses. Each
 * entry point below names the reference member it stands in for (paths relative to the reference
 * root). INTEGRATION.md shows the ~60-line C++ shim that gives these the reference's class names.
 *
 * Conventions: every function returns 0 on success or a negative ADYPT_E* code, and
 * adypt_last_error() returns a thread-local message for the last failure on the calling thread.
 * Handles are opaque; one host thread per handle at a time (the reference is single-threaded on its
 * GL-context thread). The callee copies every host input before returning (the reference's
 * Scene/WideBVH are stack locals destroyed right after OglScene::Initialize, Instance.cpp:12-33).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with ADYPT_ENODEV.
 *
 * Data layouts are the reference's GPU ABI, uploaded unchanged:
 *   node      80 B  WideBVHNode            src/BVH/WideBVH.hpp:13-26  == struct Node, traversal.glsl:1-5
 *   woop      48 B  3 x vec4 per leaf ref  src/Tracer/OglScene.cpp:93-116 == struct Woop, traversal.glsl:6
 *   triangle 100 B  Triangle               src/Util/Shape.hpp:70-88   == pathtracer.glsl:2-8
 *   material  64 B  GPUMaterial            src/Tracer/OglScene.hpp:19-28 == pathtracer.glsl:9-18
 *   ray       32 B  ox,oy,oz,tmin,dx,dy,dz,pad   (vec4 origin_tmin + vec3 dir of BVHIntersection)
 */
#ifndef ADYPT_B200_H
#define ADYPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADYPT_B200_VERSION 100

enum {
	ADYPT_OK = 0,
	ADYPT_EINVAL = -1,  /* bad argument */
	ADYPT_ENODEV = -2,  /* no usable CUDA device / device index out of range */
	ADYPT_ECUDA = -3,   /* CUDA runtime error (message has the cudaError string) */
	ADYPT_ENOMEM = -4,  /* allocation failed */
	ADYPT_EIO = -5,     /* file could not be written */
	ADYPT_ERANGE = -6,  /* value outside what the implementation supports (e.g. Sobol dimensions) */
	ADYPT_EINTERNAL = -7 /* an unexpected C++ exception was caught at the boundary (message has what()) */
};

/* where the buffers of a batch call live */
enum { ADYPT_MEM_HOST = 0, ADYPT_MEM_DEVICE = 1 };

/* OglPathTracer::ViewerTypes (src/Tracer/OglPathTracer.hpp:20) */
enum { ADYPT_VIEW_DIFFUSE = 0, ADYPT_VIEW_SPECULAR = 1, ADYPT_VIEW_EMISSIVE = 2, ADYPT_VIEW_RADIANCE = 3,
       ADYPT_VIEW_NORMAL = 4, ADYPT_VIEW_POSITION = 5 };

typedef struct adypt_scene adypt_scene;
typedef struct adypt_tracer adypt_tracer;

const char *adypt_last_error(void);
int adypt_version(void);
int adypt_device_count(int *count);

/* ------------------------------------------------------------------------------------------------
 * Scene: replaces OglScene::Initialize(const Scene&, const WideBVH&) (src/Tracer/OglScene.hpp:43,
 * OglScene.cpp:46-49,118-141), called from Instance.cpp:33.
 * nodes/tri_indices come from WideBVH::GetNodes()/GetTriIndices() (WideBVH.hpp:39-40), triangles from
 * Scene::GetTriangles() (Scene.hpp:22), materials as OglScene::init_materials builds them
 * (OglScene.cpp:51-91). woop may be NULL: the Woop rows are then built on the GPU from `triangles`
 * exactly as OglScene::init_triangles does (OglScene.cpp:93-116, glm::inverse arithmetic order).
 * triangles/materials may be NULL (n = 0) for a traversal-only scene (adypt_trace_* still work). */
typedef struct {
	int32_t device;              /* CUDA device ordinal */
	const void *nodes;           /* n_nodes * 80 B */
	uint32_t n_nodes;
	const int32_t *tri_indices;  /* n_refs, leaf order -> scene triangle id */
	uint32_t n_refs;
	const float *woop;           /* n_refs * 12 floats, or NULL */
	const void *triangles;       /* n_tris * 100 B, or NULL */
	uint32_t n_tris;
	const void *materials;       /* n_mats * 64 B, or NULL */
	uint32_t n_mats;
} adypt_scene_desc;

int adypt_scene_create(const adypt_scene_desc *desc, adypt_scene **out);
/* Diffuse textures: replaces OglScene::load_texture + the bindless handle table (OglScene.cpp:12-43, 85-90).
 * Texture i is what material.dtex == i refers to: width*height RGB8 texels, first row first (stbi_load order).
 * Sampling reproduces texture(sampler2D, uv).rgb with GL_REPEAT / GL_LINEAR / no mip-maps
 * (pathtracer.glsl:87-98, OglScene.cpp:33-38) as an explicit fp32 bilinear blend. With zero textures the scene
 * behaves like the reference compiled with TEXTURE_COUNT == 0 (dtex is ignored). Call before creating tracers. */
typedef struct {
	const uint8_t *rgb8;
	int32_t width, height;
} adypt_texture;
int adypt_scene_set_textures(adypt_scene *scene, const adypt_texture *textures, uint32_t n_textures);
int adypt_scene_destroy(adypt_scene *scene);
/* copies the device Woop array (n_refs * 12 floats) back to the host, for parity checks */
int adypt_scene_read_woop(adypt_scene *scene, float *out);
/* bytes resident on the device for this scene */
int adypt_scene_device_bytes(adypt_scene *scene, uint64_t *bytes);

/* ------------------------------------------------------------------------------------------------
 * Batch traversal: replaces void BVHIntersection(vec4 origin_tmin, vec3 dir, inout int idx, inout vec2 uv)
 * (shaders/traversal.glsl:14-255) and bool BVHIntersection(vec4, vec3) (traversal.glsl:257-494) as
 * array-in / array-out calls. rays: n * 8 floats. tri: scene triangle id or -1. t: hit distance along
 * the NORMALISED direction (1e9 on a miss; the GLSL keeps it in a local), may be NULL. uv: n * 2
 * floats, left 0 on a miss, may be NULL. memspace says whether ALL array pointers of the call are host
 * or device pointers (device pointers must belong to the scene's device). stream: a cudaStream_t
 * (NULL = default stream). Device calls are asynchronous on `stream`; host calls return when the
 * results are in the output arrays. */
int adypt_trace_closest(adypt_scene *scene, const float *rays, uint64_t n, int32_t *tri, float *t, float *uv,
                        int memspace, void *stream);
int adypt_trace_any(adypt_scene *scene, const float *rays, uint64_t n, uint8_t *occluded, int memspace, void *stream);
/* Instrumented closest-hit pass over the same rays (results discarded): out[0] = nodes visited, out[1] =
 * triangles tested, out[2] = rays that hit, out[3] = deepest traversal stack. These are the per-ray work
 * counts behind the roofline's algorithmic bytes; the oracle counts the same events. Blocking. */
int adypt_trace_stats(adypt_scene *scene, const float *rays, uint64_t n, int memspace, uint64_t out[4]);
/* number of kernel launches adypt_* calls have issued so far on this scene's device (bench accounting) */
int adypt_launch_count(uint64_t *launches);
/* tuning knobs of the persistent traversal kernel (0 = default): CTAs per SM, the refill threshold, and a
 * variant of the traversal kernels (0..22, see trace_kernel_for / launch_trace in csrc/scene.cu: 0 = the product kernel, the rest are the
 * measured alternatives kept for A/B runs, e.g. 9-11 the shared-memory hit-mask table, 12 the shared-memory ray pool; identical results) */
int adypt_trace_configure(adypt_scene *scene, int ctas_per_sm, int refill_threshold, int variant);

/* ------------------------------------------------------------------------------------------------
 * Tracer: replaces OglPathTracer (src/Tracer/OglPathTracer.hpp:66-82).
 * adypt_pt_config has the memory layout of InstanceConfig::PT (src/InstanceConfig.hpp:22-28), so a
 * `const InstanceConfig::PT *` can be passed as is. It is COPIED at create/set_config time; the reference
 * borrows the pointer and re-reads it whenever spp == 0 (OglPathTracer.cpp:39-41, 214-225) -- the C++
 * shim in INTEGRATION.md keeps that behaviour by calling adypt_tracer_set_config before the first sample. */
typedef struct {
	int32_t invocation_size; /* GL work-group edge; accepted and ignored */
	int32_t stack_size;      /* reference: unchecked GLSL array size; here: must be >= scene depth or traversal spills to a slower overflow stack, never out of bounds */
	int32_t max_bounce;      /* path segments including the primary one */
	int32_t subpixel;
	int32_t tmp_lifetime;
	float ray_tmin;
	float clamp;
	float sun[3];
} adypt_pt_config;

/* OglPathTracer::Initialize(const InstanceConfig::PT*, const OglScene&, int w, int h) (OglPathTracer.cpp:11-25).
 * bias_seed seeds the per-pixel Cranley-Patterson bias image the reference fills from std::random_device
 * (OglPathTracer.cpp:154-162); use adypt_tracer_set_bias to supply explicit bytes instead. */
int adypt_tracer_create(adypt_scene *scene, const adypt_pt_config *config, int32_t width, int32_t height,
                        uint64_t bias_seed, adypt_tracer **out);
int adypt_tracer_destroy(adypt_tracer *tracer);
int adypt_tracer_set_config(adypt_tracer *tracer, const adypt_pt_config *config);
/* width*height*2 bytes, the RG8 texels of uSobolBiasImg in row-major pixel order */
int adypt_tracer_set_bias(adypt_tracer *tracer, const uint8_t *rg8);
int adypt_tracer_get_bias(adypt_tracer *tracer, uint8_t *rg8);

/* The wavefront's CONNECT stage. The reference carries a sun-visibility test in its shader, commented out
 * (shaders/pathtracer.glsl:132): a path that leaves the scene only receives the `sun` radiance if an any-hit ray
 * from its last origin towards normalize(vec3(0.6, 1, 0.2)) is unoccluded. enabled != 0 switches that line on
 * (direction is normalised inside); the default, 0, is the reference's shipped behaviour. */
int adypt_tracer_set_sun_visibility(adypt_tracer *tracer, int32_t enabled, const float direction[3]);

/* Russian roulette, an OPT-IN extension (BASELINE.json's configs[2] asks for it; shaders/pathtracer.glsl has none, so
 * the default -- start_bounce < 0 -- is the reference's behaviour). With start_bounce >= 0, after shading bounce
 * b >= start_bounce a path survives with probability p = min(1, max(throughput.rgb)) and its throughput is divided
 * by p; the draw is Sobol dimension 2*maxBounce + b shifted by the pixel's bias like the other draws (needs
 * 3*maxBounce <= 64, else ADYPT_ERANGE). Unbiased: the expected image is unchanged. */
int adypt_tracer_set_russian_roulette(adypt_tracer *tracer, int32_t start_bounce);

/* OglPathTracer::SetCamera(const mat4& projection, const mat4& view, const vec3& position)
 * (OglPathTracer.cpp:27-32): column-major float[16]; the inverses are computed inside with glm::inverse's
 * arithmetic order. */
int adypt_tracer_set_camera(adypt_tracer *tracer, const float projection[16], const float view[16], const float position[3]);
/* Camera::GetProjection / GetView (src/Tracer/Camera.cpp:13-23) for callers without glm */
int adypt_camera_matrices(float fov_deg, float yaw_deg, float pitch_deg, int32_t width, int32_t height,
                          float projection[16], float view[16]);

/* The vector Sobol::Next writes on its (index+1)-th call after Reset(dim) (src/Util/Sobol.cpp:16-21), in closed form (the
 * state after n calls is the XOR of the direction numbers selected by the Gray code of n): what the wavefront's samplers use,
 * here for callers and tests on the host. dim <= 10005 (Sobol.hpp:9), else ADYPT_ERANGE. No GPU needed. */
int adypt_sobol_vector(uint32_t dim, uint32_t index, float *out);

/* OglPathTracer::Trace(false) (OglPathTracer.cpp:52-60): resets spp to 0 and renders the AOV `viewer_type`
 * (primaryray.glsl:46-94) into the result image. */
int adypt_tracer_primary(adypt_tracer *tracer, int32_t viewer_type);
/* n_spp x OglPathTracer::Trace(true) (OglPathTracer.cpp:36-50): the first call after create/primary clears
 * the result image and resets the Sobol generator; each sample advances spp by one. The result image is
 * the reference's running mean (pathtracer.glsl:224-226). */
int adypt_tracer_sample(adypt_tracer *tracer, int32_t n_spp);
/* Sample-sharded rendering (not in the reference): adds the clamped radiance of samples
 * first_spp .. first_spp+n_spp-1 into a SUM accumulator (w*h*4 floats, .w counts samples). first_spp must
 * be a multiple of tmp_lifetime. The accumulator is exposed as a device pointer so the caller can reduce it
 * across GPUs (NCCL) before adypt_tracer_resolve_sum divides by the sample count into the result image. */
int adypt_tracer_accumulate(adypt_tracer *tracer, int32_t first_spp, int32_t n_spp);
int adypt_tracer_sum_buffer(adypt_tracer *tracer, float **device_ptr, uint64_t *n_floats);
int adypt_tracer_clear_sum(adypt_tracer *tracer);
int adypt_tracer_resolve_sum(adypt_tracer *tracer);
/* OglPathTracer::GetSPP (OglPathTracer.hpp:76) */
int adypt_tracer_spp(adypt_tracer *tracer, int32_t *spp);
/* the result image uOutImg: width*height*4 floats RGBA, row-major, top row first (glGetTextureImage order
 * of OglPathTracer.cpp:205 is RGB; use channels = 3 for that) */
int adypt_tracer_read(adypt_tracer *tracer, float *out, int32_t channels);
int adypt_tracer_result_buffer(adypt_tracer *tracer, float **device_ptr, uint64_t *n_floats);
/* OglPathTracer::SaveResult(const char*, bool save_as_fp16) (OglPathTracer.cpp:199-212): RGB scanline
 * OpenEXR, ZIP-compressed, half or float channels */
int adypt_tracer_save_exr(adypt_tracer *tracer, const char *filename, int32_t save_as_fp16);
/* blocks until everything queued on the tracer's stream has finished */
int adypt_tracer_sync(adypt_tracer *tracer);
/* the pinhole rays of primaryray.glsl:39-44 (bias 0,0) for the current camera, written to `rays`
 * (width*height*8 floats, row-major pixel order): the C1 benchmark's ray set */
int adypt_tracer_primary_rays(adypt_tracer *tracer, float *rays, int memspace);
/* per-tracer statistics since creation: traced path segments and kernel launches */
int adypt_tracer_stats(adypt_tracer *tracer, uint64_t *segments, uint64_t *launches);

/* Measurement hooks (bench.py; not in the reference, whose only timer is a wall clock, OglPathTracer.hpp:78-82).
 * flags bit 0: a CUDA event pair around every stage launch of the wavefront -- adypt_tracer_get_profile then returns the time and
 * launch count per stage; bit 1: the traversal launches run the INSTRUMENTED kernel, which counts the nodes visited, triangles
 * tested, hits and rays of the wavefront's queues (the per-ray work behind the roofline's algorithmic bytes, SURVEY.md 8d;
 * slower, so never set in a timed run). 0 switches both off (the default). */
enum {
	ADYPT_STAGE_GENERATE = 0,      /* camera rays */
	ADYPT_STAGE_TRACE_PRIMARY = 1, /* traversal of the primary rays (once per tmpLifetime block) */
	ADYPT_STAGE_SHADE_PRIMARY = 2, /* bounce 0 of every sample from the cached primary hit */
	ADYPT_STAGE_TRACE_BOUNCE = 3,  /* traversal of the bounce queues */
	ADYPT_STAGE_SHADE_BOUNCE = 4,
	ADYPT_STAGE_ACCUMULATE = 5,
	ADYPT_STAGE_CONNECT = 6,       /* optional sun-visibility stage */
	ADYPT_STAGE_OTHER = 7,         /* Sobol vectors, queue counters */
	ADYPT_STAGE_COUNT = 8
};
typedef struct {
	double stage_ms[ADYPT_STAGE_COUNT];
	uint64_t stage_launches[ADYPT_STAGE_COUNT];
	uint64_t trace_nodes, trace_tris, trace_hits, trace_rays, trace_max_depth; /* the bounce queues (bounce >= 1) */
	uint64_t primary_nodes, primary_tris, primary_hits, primary_rays;          /* the primary rays (one traversal per tmpLifetime block) */
} adypt_tracer_profile;
int adypt_tracer_set_profiling(adypt_tracer *tracer, int32_t flags);
/* sums since the last reset; blocks until the tracer's stream is idle */
int adypt_tracer_get_profile(adypt_tracer *tracer, adypt_tracer_profile *out, int32_t reset);
/* The name (as the driver / ncu / cuobjdump show it, mangled) of the traversal kernel adypt_trace_closest (any_hit = 0) or
 * adypt_trace_any (any_hit = 1) launches for this scene's current tuning variant. */
int adypt_trace_kernel_name(adypt_scene *scene, int32_t any_hit, char *buf, uint64_t cap);

/* Evaluates the shading stage's deterministic sin/cos (op 0: out = sin(x), out2 = cos(x)) or pow (op 1: out =
 * pow(x, y)) ON THE GPU for n host values: lets tests check the CUDA copy of the recipe against the CPU oracle's
 * bit for bit. Host pointers; y / out2 may be NULL when unused. */
int adypt_debug_math(int32_t device, int32_t op, const float *x, const float *y, uint64_t n, float *out, float *out2);

/* standalone EXR writer used by adypt_tracer_save_exr (rgb: width*height*3 floats) */
int adypt_write_exr(const char *filename, const float *rgb, int32_t width, int32_t height, int32_t save_as_fp16);

/* ------------------------------------------------------------------------------------------------
 * Host side (no GPU needed): the CPU stages that FEED the tracer in the reference -- Scene (src/Util/Scene.*),
 * SBVHBuilder + WideBVHBuilder (src/BVH) and the .bvh cache (src/BVH/WideBVH.cpp) -- rebuilt from scratch
 * with byte-identical output. A host scene owns Triangle[], GPUMaterial[], the 80-byte
 * node array and the leaf-order index array. */
typedef struct adypt_host_scene adypt_host_scene;

typedef struct { /* InstanceConfig::BVH (src/InstanceConfig.hpp:15-20) */
	int32_t max_spatial_depth;
	float triangle_sah;
	float node_sah;
} adypt_bvh_config;

typedef struct {
	uint32_t n_tris, n_mats, n_nodes, n_refs, n_binary_nodes;
	const void *triangles;      /* n_tris * 100 B */
	const void *materials;      /* n_mats * 64 B */
	const void *nodes;          /* n_nodes * 80 B (NULL before a BVH is built or loaded) */
	const int32_t *tri_indices; /* n_refs */
	const void *binary_nodes;   /* n_binary_nodes * 32 B, the intermediate SBVH (NULL when loaded from a .bvh file) */
	float aabb[6];              /* Scene::GetAABB: min xyz, max xyz */
} adypt_host_scene_info;

/* Scene::LoadFromFile (src/Util/Scene.cpp:9-136): Wavefront OBJ + MTL -> Triangle[] (flat normals generated when
 * the file has none, v texture coordinate flipped) and materials as OglScene::init_materials lays them out. */
int adypt_host_scene_load_obj(const char *obj_path, adypt_host_scene **out);
/* the same Triangle records from raw corner positions (n_tris * 9 floats) + material ids, with the flat normals
 * Scene.cpp:117-123 generates: for callers that already hold the mesh in memory */
int adypt_host_scene_from_triangles(const float *positions, const int32_t *material_ids, uint32_t n_tris,
                                    const void *materials64, uint32_t n_mats, adypt_host_scene **out);
int adypt_host_scene_destroy(adypt_host_scene *scene);
/* SBVHBuilder::Run + WideBVHBuilder::Run (Instance.cpp:22-24) */
int adypt_host_scene_build_bvh(adypt_host_scene *scene, const adypt_bvh_config *config);
/* WideBVH::LoadFromFile / SaveToFile (src/BVH/WideBVH.cpp:9-66). load returns ADYPT_EIO when the file is missing,
 * has another magic or was built with other parameters -- the caller then rebuilds, as Instance.cpp:20-31 does */
int adypt_host_scene_load_bvh(adypt_host_scene *scene, const char *bvh_path, const adypt_bvh_config *expected);
int adypt_host_scene_save_bvh(adypt_host_scene *scene, const char *bvh_path, const adypt_bvh_config *config);
int adypt_host_scene_get(adypt_host_scene *scene, adypt_host_scene_info *info);
/* Decodes the map_Kd files the OBJ's materials name (PNG, TGA) and renumbers Material::dtex the way
 * OglScene::init_materials does: index among the textures that loaded, -1 when loading failed. Returns the
 * number loaded / failed through the out-parameters (either may be NULL). adypt_host_scene_upload sends them along. */
int adypt_host_scene_load_textures(adypt_host_scene *scene, uint32_t *n_loaded, uint32_t *n_failed);
/* texture i of the host scene after adypt_host_scene_load_textures (RGB8, first row first) */
int adypt_host_scene_texture(adypt_host_scene *scene, uint32_t i, const uint8_t **rgb8, int32_t *width, int32_t *height);
/* OglScene::Initialize(scene, wbvh) (Instance.cpp:33): uploads to `device`, Woop rows built on the GPU */
int adypt_host_scene_upload(adypt_host_scene *scene, int32_t device, adypt_scene **out);

/* ------------------------------------------------------------------------------------------------
 * Render group: one process driving several GPUs of a box (not in the reference, which is single-GPU).
 * Sample-index sharding as in SURVEY.md 8e: blocks of tmpLifetime samples round-robin over the devices, scene
 * replicated, ONE ncclReduce of the W*H*4-float sum accumulator to the first device, which resolves the image.
 * NCCL is loaded with dlopen on first use; a one-device group does not need it. The per-process alternative
 * (one rank per GPU, torch.distributed) is tools/render_sharded.py. */
typedef struct adypt_group adypt_group;
int adypt_group_create(adypt_host_scene *scene, const adypt_pt_config *config, int32_t width, int32_t height, uint64_t bias_seed,
                       const int32_t *devices, uint32_t n_devices, adypt_group **out);
int adypt_group_destroy(adypt_group *group);
int adypt_group_set_camera(adypt_group *group, const float projection[16], const float view[16], const float position[3]);
int adypt_group_set_sun_visibility(adypt_group *group, int32_t enabled, const float direction[3]);
int adypt_group_set_russian_roulette(adypt_group *group, int32_t start_bounce);
int adypt_group_render(adypt_group *group, int32_t total_spp); /* samples [0, total_spp); blocks until the image is resolved */
int adypt_group_read(adypt_group *group, float *out, int32_t channels);
int adypt_group_save_exr(adypt_group *group, const char *filename, int32_t save_as_fp16);
/* the CUDA stream a tracer enqueues its work on (a cudaStream_t), for callers that add their own device work */
int adypt_tracer_stream(adypt_tracer *tracer, void **stream);

/* ------------------------------------------------------------------------------------------------
 * The .config instance file: InstanceConfig (src/InstanceConfig.hpp:12-48). Same JSON schema, the same
 * acceptance rules (every key mandatory; "Float" values must be written as doubles -- 45.0, not 45 -- exactly
 * as rapidjson's IsFloat demands, InstanceConfig.cpp:17-18) and the same PrettyWriter text on output. */
typedef struct {
	int32_t width, height;
	adypt_bvh_config bvh;
	adypt_pt_config pt;
	struct {
		float speed, mouse_sensitive, fov, yaw, pitch;
		float position[3];
	} cam;                    /* InstanceConfig::Cam (InstanceConfig.hpp:30-35) */
	char obj_filename[1024];  /* scene.filename */
	char bvh_filename[1024];  /* bvh.filename */
} adypt_instance_config;

int adypt_config_set_default(adypt_instance_config *config);                 /* InstanceConfig::SetDefault */
int adypt_config_load(const char *path, adypt_instance_config *config);      /* InstanceConfig::LoadFromFile; the
                                                 "[PARSER]ERR: ..." line the reference prints is in adypt_last_error() */
int adypt_config_to_json(const adypt_instance_config *config, char *buf, uint64_t cap, uint64_t *needed); /* GetJson */
int adypt_config_save(const adypt_instance_config *config, const char *path); /* InstanceConfig::SaveToFile */
/* The text a finite double gets in a saved .config: rapidjson's Writer::WriteDouble (Grisu2 digits, then Prettify).
 * out must hold 32 bytes; NaN / infinity give ADYPT_EINVAL (rapidjson's writer refuses them too). */
int adypt_config_format_double(double value, char out[32]);

#ifdef __cplusplus
}
#endif
#endif /* ADYPT_B200_H */
