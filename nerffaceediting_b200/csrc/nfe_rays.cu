// Ray generation, ray/box limits and stratified coarse depths (one thread per ray / per depth).
// Replaces RaySampler.forward (ray_sampler.py:24-63, ~15 tiny ATen launches), get_ray_limits_box
// (math_utils.py:46-98) and sample_stratified (renderer.py:169-192).
#include "nfe_common.cuh"

namespace nfe {

__global__ void generate_rays_kernel(const float* __restrict__ cam2world, const float* __restrict__ intrinsics, int n, int res,
                                     float* __restrict__ origins, float* __restrict__ dirs)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t rays = (int64_t)res * res;
    if (idx >= (int64_t)n * rays) return;
    const int b = (int)(idx / rays);
    const int m = (int)(idx % rays);
    const int i = m / res, j = m % res;
    const float* c = cam2world + 16 * b;
    const float* k = intrinsics + 9 * b;
    const float fx = k[0], sk = k[1], cx = k[2], fy = k[4], cy = k[5];
    const float inv_res = __fdiv_rn(1.0f, (float)res), half = __fdiv_rn(0.5f, (float)res);
    const float x_cam = __fadd_rn(__fmul_rn((float)j, inv_res), half);
    const float y_cam = __fadd_rn(__fmul_rn((float)i, inv_res), half);
    // x_lift = (x_cam - cx + cy*sk/fy - sk*y_cam/fy) / fx ; y_lift = (y_cam - cy) / fy   (ray_sampler.py:51-52)
    const float t0 = __fsub_rn(x_cam, cx);
    const float t1 = __fdiv_rn(__fmul_rn(cy, sk), fy);
    const float t2 = __fdiv_rn(__fmul_rn(sk, y_cam), fy);
    const float x_lift = __fdiv_rn(__fsub_rn(__fadd_rn(t0, t1), t2), fx);
    const float y_lift = __fdiv_rn(__fsub_rn(y_cam, cy), fy);
    float d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(c[4 * r + 0], x_lift);
        acc = __fadd_rn(acc, __fmul_rn(c[4 * r + 1], y_lift));
        acc = __fadd_rn(acc, c[4 * r + 2]);
        acc = __fadd_rn(acc, c[4 * r + 3]);
        d[r] = __fsub_rn(acc, c[4 * r + 3]);
    }
    float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    nrm = fmaxf(nrm, 1e-12f);  // F.normalize eps
    float* o = origins + idx * 3;
    float* dd = dirs + idx * 3;
    o[0] = c[3]; o[1] = c[7]; o[2] = c[11];
    dd[0] = __fdiv_rn(d[0], nrm); dd[1] = __fdiv_rn(d[1], nrm); dd[2] = __fdiv_rn(d[2], nrm);
}

__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }

__global__ void ray_limits_box_kernel(const float* __restrict__ origins, const float* __restrict__ dirs, int64_t n_rays, float lo, float hi,
                                      float* __restrict__ tmin_out, float* __restrict__ tmax_out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float* o = origins + 3 * r;
    const float* d = dirs + 3 * r;
    float inv[3]; bool sg[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { inv[a] = __fdiv_rn(1.0f, d[a]); sg[a] = inv[a] < 0.0f; }
    bool valid = true;
    float tmin = __fmul_rn(__fsub_rn(sg[0] ? hi : lo, o[0]), inv[0]);
    float tmax = __fmul_rn(__fsub_rn(sg[0] ? lo : hi, o[0]), inv[0]);
    const float tymin = __fmul_rn(__fsub_rn(sg[1] ? hi : lo, o[1]), inv[1]);
    const float tymax = __fmul_rn(__fsub_rn(sg[1] ? lo : hi, o[1]), inv[1]);
    if (tmin > tymax || tymin > tmax) valid = false;
    tmin = max_nan(tmin, tymin);
    tmax = min_nan(tmax, tymax);
    const float tzmin = __fmul_rn(__fsub_rn(sg[2] ? hi : lo, o[2]), inv[2]);
    const float tzmax = __fmul_rn(__fsub_rn(sg[2] ? lo : hi, o[2]), inv[2]);
    if (tmin > tzmax || tzmin > tmax) valid = false;
    tmin = max_nan(tmin, tzmin);
    tmax = min_nan(tmax, tzmax);
    if (!valid) { tmin = -1.0f; tmax = -2.0f; }
    tmin_out[r] = tmin;
    tmax_out[r] = tmax;
}

__global__ void sample_stratified_kernel(int64_t total, int s_c, int mode, const float* __restrict__ table, float delta_scalar,
                                         float inv_start, float inv_end, const float* __restrict__ start_per_ray,
                                         const float* __restrict__ end_per_ray, const float* __restrict__ jitter, int stochastic,
                                         uint64_t seed, uint64_t offset, float* __restrict__ depths)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t r = idx / s_c;
    const int s = (int)(idx % s_c);
    float u = 0.0f;
    if (jitter) u = jitter[idx];
    else if (stochastic) u = u01(philox4x32(seed, (uint64_t)idx, offset).x);
    float t;
    if (mode == 0) {
        t = __fadd_rn(table[s], __fmul_rn(u, delta_scalar));
    } else if (mode == 1) {
        const float a = start_per_ray[r], b = end_per_ray[r];
        const float step = __fdiv_rn((float)s, (float)(s_c - 1));
        t = __fadd_rn(a, __fmul_rn(step, __fsub_rn(b, a)));
        const float delta = __fdiv_rn(__fsub_rn(b, a), (float)(s_c - 1));
        t = __fadd_rn(t, __fmul_rn(u, delta));
    } else {
        const float sp = __fadd_rn(table[s], __fmul_rn(u, delta_scalar));
        t = __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(inv_start, __fsub_rn(1.0f, sp)), __fmul_rn(inv_end, sp)));
    }
    depths[idx] = t;
}

}  // namespace nfe

using namespace nfe;

NFE_EXPORT int nfe_generate_rays(const float* cam2world, const float* intrinsics, int n, int resolution, float* origins, float* dirs,
                                 nfe_stream_t stream)
{
    if (n == 0) return 0;
    NFE_REQUIRE(cam2world && intrinsics && origins && dirs, "nfe_generate_rays: null pointer");
    NFE_REQUIRE(n >= 0 && resolution >= 1 && resolution <= 16384, "nfe_generate_rays: bad sizes");
    const int64_t total = (int64_t)n * resolution * resolution;
    if (total == 0) return 0;
    generate_rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(cam2world, intrinsics, n, resolution, origins, dirs);
    NFE_LAUNCH_CHECK("generate_rays_kernel");
    return 0;
}

NFE_EXPORT int nfe_ray_limits_box(const float* origins, const float* dirs, int64_t n_rays, float box_side_length, float* tmin, float* tmax,
                                  nfe_stream_t stream)
{
    if (n_rays <= 0) return 0;
    NFE_REQUIRE(origins && dirs && tmin && tmax, "nfe_ray_limits_box: null pointer");
    if (n_rays <= 0) return 0;
    const float hi = 1.0f * (box_side_length / 2.0f), lo = -1.0f * (box_side_length / 2.0f);
    ray_limits_box_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, as_stream(stream)>>>(origins, dirs, n_rays, lo, hi, tmin, tmax);
    NFE_LAUNCH_CHECK("ray_limits_box_kernel");
    return 0;
}

NFE_EXPORT int nfe_sample_stratified(int64_t n_rays, int s_c, int mode, const float* table, double ray_start, double ray_end,
                                     const float* start_per_ray, const float* end_per_ray, const float* jitter, int stochastic,
                                     uint64_t seed, uint64_t offset, float* depths, nfe_stream_t stream)
{
    if (n_rays <= 0) return 0;
    NFE_REQUIRE(depths, "nfe_sample_stratified: null output");
    NFE_REQUIRE(s_c >= 2, "nfe_sample_stratified: depth_resolution must be >= 2 (got %d)", s_c);
    NFE_REQUIRE(mode >= 0 && mode <= 2, "nfe_sample_stratified: bad mode %d", mode);
    NFE_REQUIRE(mode == 1 ? (start_per_ray && end_per_ray) : (table != nullptr), "nfe_sample_stratified: missing table / per-ray limits");
    if (n_rays <= 0) return 0;
    // Python-double scalars folded to fp32 exactly as the reference's tensor*scalar ops do
    const float delta = mode == 0 ? (float)((ray_end - ray_start) / (double)(s_c - 1)) : (float)(1.0 / (double)(s_c - 1));
    const float inv_start = (float)(1.0 / ray_start), inv_end = (float)(1.0 / ray_end);
    const int64_t total = n_rays * s_c;
    sample_stratified_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        total, s_c, mode, table, delta, inv_start, inv_end, start_per_ray, end_per_ray, jitter, stochastic, seed, offset, depths);
    NFE_LAUNCH_CHECK("sample_stratified_kernel");
    return 0;
}
