// Measurement aids (no product path calls these).
//
// nfe_bench_l2_gather: the L2 -> SM gather roof for the field kernel's access shape (VERDICT r01 "next" #3): random 128-byte
// texel lines of an L2-resident table, LDG.128 with 8 lanes per line = 4 lines per warp instruction, `depth` independent
// loads in flight per warp.  bench.py times it with CUDA events on the box it runs on and reports the result as
// roofline.l2_gbs_measured; profiles/microbench/l2_gather.cu is the full sweep (warps x depth, cp.async, TMA gather4).
#include "nfe_common.cuh"

namespace nfe {

__device__ __forceinline__ float4 ld_cg_v4(const float* a)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
    return v;
}

template <int DEPTH>
__global__ void __launch_bounds__(1024, 1) l2_gather_kernel(const float* __restrict__ table, uint32_t n_lines, int iters, float* sink)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    uint32_t s = (warp_id * 4 + g) * 2654435761u + 12345u;           // one index stream per lane group (= per line)
    const float* lane_base = table + c * 4;
    float acc = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float4 v[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            s = s * 1664525u + 1013904223u;
            v[d] = ld_cg_v4(lane_base + (size_t)__umulhi(s, n_lines) * 32);
        }
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) acc += (v[d].x + v[d].y) + (v[d].z + v[d].w);
    }
    if (acc == 123.456f) sink[0] = acc;       // keeps the loads alive
}

}  // namespace nfe

using namespace nfe;

NFE_EXPORT int nfe_bench_l2_gather(const float* table, int64_t n_lines, int warps_per_cta, int depth, int iters, int64_t* host_lines_out,
                                   float* sink, nfe_stream_t stream)
{
    NFE_REQUIRE(table && sink && n_lines > 0 && n_lines < (1ll << 32), "nfe_bench_l2_gather: bad table");
    NFE_REQUIRE(warps_per_cta >= 1 && warps_per_cta <= 32 && iters >= 1, "nfe_bench_l2_gather: bad launch shape");
    NFE_REQUIRE(depth == 4 || depth == 8 || depth == 12, "nfe_bench_l2_gather: depth must be 4, 8 or 12");
    const int grid = sm_count();
    cudaStream_t st = as_stream(stream);
    if (depth == 4) l2_gather_kernel<4><<<grid, warps_per_cta * 32, 0, st>>>(table, (uint32_t)n_lines, iters, sink);
    else if (depth == 8) l2_gather_kernel<8><<<grid, warps_per_cta * 32, 0, st>>>(table, (uint32_t)n_lines, iters, sink);
    else l2_gather_kernel<12><<<grid, warps_per_cta * 32, 0, st>>>(table, (uint32_t)n_lines, iters, sink);
    if (host_lines_out) *host_lines_out = (int64_t)grid * warps_per_cta * iters * depth * 4;
    NFE_LAUNCH_CHECK("l2_gather_kernel");
    return 0;
}
