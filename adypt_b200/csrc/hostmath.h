// Host-side arithmetic the reference does on the CPU before/around each dispatch.
#pragma once
#include <stdint.h>

namespace adypt {

// glm::inverse(mat4) (dep/glm/detail/func_matrix.inl:294-351), column-major float[16]
void mat4_inverse(const float in[16], float out[16]);
// Camera::GetView / Camera::GetProjection (src/Tracer/Camera.cpp:13-23), column-major float[16]
void camera_matrices(float fov_deg, float yaw_deg, float pitch_deg, int width, int height, float proj[16], float view[16]);

// Sobol direction numbers for the first sobol_max_dim() dimensions, regenerated from Joe-Kuo parameters
// (sobol_params.inc); v[j][k] is the reference's kMatrices[j][k] (src/Util/Sobol.inl)
int sobol_max_dim();
const uint32_t *sobol_directions(); // [sobol_max_dim()][32]
// the vector Sobol::Next writes on its (index+1)-th call after Reset (src/Util/Sobol.cpp:16-21)
void sobol_vector(uint32_t dim, uint32_t index, float *out);

// deterministic stand-in for the std::random_device bytes of OglPathTracer.cpp:154-162
void fill_bias(uint64_t seed, uint64_t n_bytes, uint8_t *out);

} // namespace adypt
