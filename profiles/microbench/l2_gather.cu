// L2 gather micro-benchmark for the tri-plane texel access shape (VERDICT r01 "Next round" items 2(ii) and 3).
//
// A channel-last plane set is a table of 128-byte texel lines (32 fp32 channels).  One bilinear sample reads 4 lines per
// plane: (x,y), (x+1,y) adjacent in memory, (x,y+1), (x+1,y+1) one image row further.  This program measures how many such
// lines per second one B200 delivers from an L2-resident table into the SMs
//
//   mode ldg     LDG.128, 8 lanes per line, 4 lines per warp instruction, DEPTH instructions in flight per warp
//                (the shipped field kernel's access: DEPTH = 12, 8 warps per SM)
//   mode tma     cp.async.bulk.tensor.2d.tile::gather4 (4 arbitrary 128-byte rows per instruction) into a shared-memory
//                ring, completion on mbarriers; ISSUERS threads issue, one consumer warp optionally reads the ring back
//
// over the number of warps / issuing threads and the requests in flight, for a `random` pattern (every line independent) and
// the `quad` pattern (2x2 texel footprints at random positions of a 256x256 plane, three planes per sample).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/microbench/_bin/l2_gather profiles/microbench/l2_gather.cu
//   profiles/microbench/_bin/l2_gather            # prints one JSON line per configuration
//
// Numbers land in profiles/l2_gather_r02.txt; the LDG variant also lives in the library as nfe_bench_l2_gather so that
// bench.py can report roofline.l2_gbs_measured on the box it runs on.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

struct Pattern {
    uint32_t n_lines;      // table size in 128-byte lines
    int quad;              // 0: every line independent; 1: 2x2 footprints in HxW planes
    uint32_t H, W, planes; // quad pattern geometry: planes x H x W lines
};

// line indices of one "tap quad": 4 lines (TMA issuers: one thread needs all four)
template <bool QUAD>
__device__ __forceinline__ void next_quad(const Pattern& p, uint32_t& s, uint32_t (&idx)[4])
{
    if (!QUAD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) idx[i] = __umulhi(lcg(s), p.n_lines);
    } else {
        const uint32_t base = __umulhi(lcg(s), p.n_lines - p.W - 1);
        idx[0] = base; idx[1] = base + 1; idx[2] = base + p.W; idx[3] = base + p.W + 1;
    }
}
// the line of lane group g (0..3) of the next quad, branch-free: in the random pattern every group runs its own stream, in the
// quad pattern the warp shares one stream and the group picks its corner of the 2x2 footprint
template <bool QUAD>
__device__ __forceinline__ uint32_t next_line(const Pattern& p, uint32_t& s, uint32_t corner)
{
    if (!QUAD) return __umulhi(lcg(s), p.n_lines);
    return __umulhi(lcg(s), p.n_lines - p.W - 1) + corner;
}

// ------------------------------------------------------------------------------------------------ LDG
// volatile asm keeps program order among these statements: every load of a batch is issued before the first consumer, which
// ptxas otherwise interleaves (it recycled six register quads in the first version of this file: effective depth 6)
__device__ __forceinline__ float4 ld_cg_v4(const float* a)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
    return v;
}
__device__ __forceinline__ float4 ld_nc_v4(const float* a)
{
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
    return v;
}
__device__ __forceinline__ float2 ld_cg_v2(const float* a)
{
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(a));
    return v;
}
__device__ __forceinline__ float ld_cg_f(const float* a)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(a));
    return v;
}
__device__ __forceinline__ void consume(float& acc, const float4& v)
{
    asm volatile("add.f32 %0, %0, %1;\n\tadd.f32 %0, %0, %2;\n\tadd.f32 %0, %0, %3;\n\tadd.f32 %0, %0, %4;" : "+f"(acc) : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
// the 8 lanes of a group read one line; a warp instruction reads the 4 lines of one quad (lane>>3 selects the line).
// In the field kernel the 4 groups are 4 SAMPLES and one instruction is one TAP of each; for bandwidth it is the same
// 4 lines per instruction.
// LPI = lines per warp instruction: 4 -> LDG.128 by 8 lanes per line, 2 -> LDG.64 by 16 lanes, 1 -> LDG.32 by all 32 lanes
template <int DEPTH, bool CG, bool QUAD, int LPI, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) ldg_kernel(const float* __restrict__ table, Pattern p, int iters, float* sink)
{
    constexpr int LANES = 32 / LPI;         // lanes per line; floats per lane = 32 / LANES
    constexpr int FL = 32 / LANES;
    const int lane = threadIdx.x & 31, g = lane / LANES, c = lane % LANES;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    uint32_t s = (QUAD ? warp_id : warp_id * 4 + g) * 2654435761u + 12345u;
    const uint32_t corner = (g & 1) + (g >> 1) * p.W;
    const float* lane_base = table + c * FL;
    float acc = 0.f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float4 v[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const uint32_t line = next_line<QUAD>(p, s, corner);
            const float* a = lane_base + (size_t)line * 32;
            if (FL == 4) v[d] = CG ? ld_cg_v4(a) : ld_nc_v4(a);
            else if (FL == 2) { const float2 t = ld_cg_v2(a); v[d] = make_float4(t.x, t.y, 0.f, 0.f); }
            else { const float t = ld_cg_f(a); v[d] = make_float4(t, 0.f, 0.f, 0.f); }
        }
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) consume(acc, v[d]);
    }
    if (acc == 123.456f) sink[0] = acc;
}

// rolling variant (the field kernel's pipeline): DEPTH loads stay in flight; each step consumes the oldest and re-issues it
template <int DEPTH, bool QUAD, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) ldg_rolling_kernel(const float* __restrict__ table, Pattern p, int iters, float* sink)
{
    const int lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    uint32_t s = (QUAD ? warp_id : warp_id * 4 + g) * 2654435761u + 12345u;
    const uint32_t corner = (g & 1) + (g >> 1) * p.W;
    const float* lane_base = table + c * 4;
    float acc = 0.f;
    float4 v[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) v[d] = ld_cg_v4(lane_base + (size_t)next_line<QUAD>(p, s, corner) * 32);
#pragma unroll 1
    for (int it = 1; it < iters; ++it) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            consume(acc, v[d]);
            v[d] = ld_cg_v4(lane_base + (size_t)next_line<QUAD>(p, s, corner) * 32);
        }
    }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) consume(acc, v[d]);
    if (acc == 123.456f) sink[0] = acc;
}

// cp.async (LDGSTS) variant: 16 bytes per lane straight into shared memory, DEPTH commit groups in flight, no data registers
template <int DEPTH, bool QUAD>
__global__ void cpasync_kernel(const float* __restrict__ table, Pattern p, int iters, float* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, c = lane & 7;
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    uint32_t s = (QUAD ? warp_id : warp_id * 4 + g) * 2654435761u + 12345u;
    const uint32_t corner = (g & 1) + (g >> 1) * p.W;
    const float* lane_base = table + c * 4;
    const uint32_t dst0 = smem_u32(smem) + (warp * DEPTH * 32 + lane) * 16;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const float* a = lane_base + (size_t)next_line<QUAD>(p, s, corner) * 32;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst0 + d * 512), "l"(a) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const float4 t = *reinterpret_cast<const float4*>(smem + (warp * DEPTH * 32 + lane) * 16);
        acc += t.x;
    }
    if (acc == 123.456f) sink[0] = acc;
}

// ------------------------------------------------------------------------------------------------ TMA gather4
__device__ __forceinline__ void mbar_init(uint64_t* m, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(m)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* m, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(m)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* m) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(m)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* m, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra W_DONE;\n\tbra W_LOOP;\n\tW_DONE:\n\t}\n"
        :: "r"(smem_u32(m)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tma_gather4(const CUtensorMap* tmap, uint64_t* mbar, void* dst, int col, int r0, int r1, int r2, int r3)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :: "r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(mbar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// Ring of STAGES stages; a stage holds QUADS gather4 results (QUADS x 512 bytes).  ISSUERS threads of warp 0.. fill a stage
// (each issues QUADS / ISSUERS instructions); consumer warps wait for the stage, optionally read it (LDS.128, like a blend
// would), and release it.
constexpr int TMA_QUADS = 48;           // 48 quads = 24 KB per stage = 16 samples x 3 planes
template <int STAGES, bool QUAD>
__global__ void tma_kernel(const __grid_constant__ CUtensorMap tmap, Pattern p, int iters, int issuers, int consume, float* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + STAGES;
    unsigned char* ring = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_issue_warps = (issuers + 31) / 32;
    const int n_cons_warps = (blockDim.x >> 5) - n_issue_warps;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], n_cons_warps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    if (warp < n_issue_warps) {
        const int me = threadIdx.x;
        uint32_t s = (blockIdx.x * 1024 + me) * 2654435761u + 777u;
        for (int it = 0; it < iters; ++it) {
            const int st = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            if (it >= STAGES) mbar_wait(&empty[st], ph ^ 1);
            if (me == 0) mbar_expect_tx(&full[st], TMA_QUADS * 512);
            __syncwarp();
            if (me < issuers) {
                for (int q = me; q < TMA_QUADS; q += issuers) {
                    uint32_t idx[4];
                    next_quad<QUAD>(p, s, idx);
                    tma_gather4(&tmap, &full[st], ring + (size_t)st * TMA_QUADS * 512 + q * 512, 0, (int)idx[0], (int)idx[1], (int)idx[2], (int)idx[3]);
                }
            }
        }
    } else {
        const int cw = warp - n_issue_warps;
        for (int it = 0; it < iters; ++it) {
            const int st = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full[st], ph);
            if (consume) {
                const float4* src = reinterpret_cast<const float4*>(ring + (size_t)st * TMA_QUADS * 512);
                for (int i = cw * 32 + lane; i < TMA_QUADS * 32; i += n_cons_warps * 32) {
                    const float4 v = src[i];
                    acc += v.x + v.y + v.z + v.w;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g_sms = 148;
static float* g_sink = nullptr;

template <class Launch>
static float best_ms(Launch launch)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();                                  // warm-up: table into L2
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = fminf(best, ms);
    }
    CK(cudaGetLastError());
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return best;
}

static void report(const char* mode, const char* pat, double mb, int warps, int ctas, int depth, int lpi, float ms, double lines)
{
    printf("{\"mode\": \"%s\", \"pattern\": \"%s\", \"table_mb\": %.1f, \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"depth\": %d, \"lines_per_instr\": %d, "
           "\"ms\": %.4f, \"gbs\": %.1f, \"lines_per_us_per_sm\": %.1f}\n", mode, pat, mb, warps, ctas, depth, lpi, ms, lines * 128 / ms * 1e-6,
           lines / ms * 1e-3 / g_sms);
    fflush(stdout);
}

template <int DEPTH, bool CG, bool QUAD, int LPI>
static void run_ldg(const float* table, Pattern p, int warps, int ctas, double mb)
{
    const int iters = 6144 / DEPTH;
    const int grid = g_sms * ctas;
    const float ms = best_ms([&] {
        if (warps <= 8 && ctas == 1) ldg_kernel<DEPTH, CG, QUAD, LPI, 256><<<grid, warps * 32>>>(table, p, iters, g_sink);
        else if (warps <= 16 && ctas == 1) ldg_kernel<DEPTH, CG, QUAD, LPI, 512><<<grid, warps * 32>>>(table, p, iters, g_sink);
        else ldg_kernel<DEPTH, CG, QUAD, LPI, 1024><<<grid, warps * 32>>>(table, p, iters, g_sink);
    });
    report(CG ? "ldg.cg" : "ldg.nc", QUAD ? "quad" : "random", mb, warps, ctas, DEPTH, LPI, ms, (double)grid * warps * iters * DEPTH * LPI);
}
template <int DEPTH, bool QUAD>
static void run_rolling(const float* table, Pattern p, int warps, int ctas, double mb)
{
    const int iters = 6144 / DEPTH;
    const int grid = g_sms * ctas;
    const float ms = best_ms([&] {
        if (warps <= 8) ldg_rolling_kernel<DEPTH, QUAD, 256><<<grid, warps * 32>>>(table, p, iters, g_sink);
        else ldg_rolling_kernel<DEPTH, QUAD, 512><<<grid, warps * 32>>>(table, p, iters, g_sink);
    });
    report("ldg.cg.rolling", QUAD ? "quad" : "random", mb, warps, ctas, DEPTH, 4, ms, (double)grid * warps * iters * DEPTH * 4);
}
template <int DEPTH, bool QUAD>
static void run_cpasync(const float* table, Pattern p, int warps, int ctas, double mb)
{
    const int iters = 6144 / DEPTH;
    const int grid = g_sms * ctas;
    const size_t smem = (size_t)warps * DEPTH * 512;
    CK(cudaFuncSetAttribute(cpasync_kernel<DEPTH, QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float ms = best_ms([&] { cpasync_kernel<DEPTH, QUAD><<<grid, warps * 32, smem>>>(table, p, iters, g_sink); });
    report("cp.async", QUAD ? "quad" : "random", mb, warps, ctas, DEPTH, 4, ms, (double)grid * warps * iters * DEPTH * 4);
}
template <int STAGES, bool QUAD>
static void run_tma(const CUtensorMap& tmap, Pattern p, int issuers, int cons_warps, int consume, double mb)
{
    const size_t smem = 1024 + (size_t)STAGES * TMA_QUADS * 512;
    CK(cudaFuncSetAttribute(tma_kernel<STAGES, QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 256;
    const int threads = ((issuers + 31) / 32 + cons_warps) * 32;
    const float ms = best_ms([&] { tma_kernel<STAGES, QUAD><<<g_sms, threads, smem>>>(tmap, p, iters, issuers, consume, g_sink); });
    const double lines = (double)g_sms * iters * TMA_QUADS * 4;
    printf("{\"mode\": \"tma_gather4\", \"pattern\": \"%s\", \"table_mb\": %.1f, \"stages\": %d, \"stage_kb\": %d, \"issuers\": %d, \"consumer_warps\": %d, "
           "\"consume\": %d, \"ms\": %.4f, \"gbs\": %.1f, \"lines_per_us_per_sm\": %.1f}\n", QUAD ? "quad" : "random", mb, STAGES, TMA_QUADS / 2, issuers,
           cons_warps, consume, ms, lines * 128 / ms * 1e-6, lines / ms * 1e-3 / g_sms);
    fflush(stdout);
}

template <bool QUAD>
static void sweep(const float* table, Pattern p, double mb, const CUtensorMap& tmap, bool full)
{
    // the shipped field kernel's shape first: 8 warps, 12 LDG.128 (4 lines each) in flight
    run_ldg<12, true, QUAD, 4>(table, p, 8, 1, mb);
    run_rolling<12, QUAD>(table, p, 8, 1, mb);
    if (!full) { run_ldg<4, true, QUAD, 4>(table, p, 32, 2, mb); return; }
    run_ldg<12, false, QUAD, 4>(table, p, 8, 1, mb);
    // depth at 8 warps
    run_ldg<1, true, QUAD, 4>(table, p, 8, 1, mb);
    run_ldg<2, true, QUAD, 4>(table, p, 8, 1, mb);
    run_ldg<4, true, QUAD, 4>(table, p, 8, 1, mb);
    run_ldg<8, true, QUAD, 4>(table, p, 8, 1, mb);
    run_ldg<24, true, QUAD, 4>(table, p, 8, 1, mb);
    // warps at depth 4 / 8 / 12
    for (int w : {4, 12, 16, 20, 24, 32}) run_ldg<4, true, QUAD, 4>(table, p, w, 1, mb);
    for (int w : {4, 12, 16, 20, 24, 32}) run_ldg<8, true, QUAD, 4>(table, p, w, 1, mb);
    for (int w : {16, 32}) run_ldg<12, true, QUAD, 4>(table, p, w, 1, mb);
    run_ldg<2, true, QUAD, 4>(table, p, 32, 2, mb);
    run_ldg<4, true, QUAD, 4>(table, p, 32, 2, mb);
    run_ldg<8, true, QUAD, 4>(table, p, 32, 2, mb);
    // lines per instruction
    run_ldg<12, true, QUAD, 2>(table, p, 8, 1, mb);
    run_ldg<12, true, QUAD, 1>(table, p, 8, 1, mb);
    run_ldg<24, true, QUAD, 1>(table, p, 8, 1, mb);
    run_ldg<12, true, QUAD, 1>(table, p, 32, 1, mb);
    run_ldg<8, true, QUAD, 1>(table, p, 32, 2, mb);
    // rolling, cp.async
    run_rolling<4, QUAD>(table, p, 8, 1, mb);
    run_rolling<12, QUAD>(table, p, 16, 1, mb);
    run_rolling<4, QUAD>(table, p, 16, 1, mb);
    run_cpasync<4, QUAD>(table, p, 8, 1, mb);
    run_cpasync<12, QUAD>(table, p, 8, 1, mb);
    run_cpasync<24, QUAD>(table, p, 8, 1, mb);
    run_cpasync<12, QUAD>(table, p, 16, 1, mb);
    run_cpasync<12, QUAD>(table, p, 32, 1, mb);
    // TMA gather4
    run_tma<4, QUAD>(tmap, p, 1, 1, 0, mb);
    run_tma<4, QUAD>(tmap, p, 8, 1, 0, mb);
    run_tma<4, QUAD>(tmap, p, 32, 1, 0, mb);
    run_tma<4, QUAD>(tmap, p, 64, 1, 0, mb);
    run_tma<4, QUAD>(tmap, p, 128, 1, 0, mb);
    run_tma<8, QUAD>(tmap, p, 128, 1, 0, mb);
    run_tma<8, QUAD>(tmap, p, 256, 1, 0, mb);
    run_tma<8, QUAD>(tmap, p, 4 * 32 - 31, 1, 0, mb);      // 4 warps, one issuing lane each (lanes 0, 32, 64, 96 -> issuers stride)
    run_tma<8, QUAD>(tmap, p, 128, 8, 1, mb);
}

int main(int argc, char** argv)
{
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    CK(cudaMalloc(&g_sink, 4));
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { fprintf(stderr, "cuTensorMapEncodeTiled not available\n"); return 1; }

    // tables: one plane set of 3 x 256 x 256 texels (25.2 MB, L2-resident), and 8 items (201 MB, larger than L2)
    for (double items : {1.0, 8.0}) {
        const uint32_t H = 256, W = 256, planes = (uint32_t)(3 * items);
        const uint32_t n_lines = planes * H * W;
        const double mb = n_lines * 128.0 / 1e6;
        float* table;
        CK(cudaMalloc(&table, (size_t)n_lines * 128));
        CK(cudaMemset(table, 0, (size_t)n_lines * 128));
        CUtensorMap tmap;
        const cuuint64_t gdim[2] = {32, n_lines};            // 32 fp32 per row, one row per texel line
        const cuuint64_t gstride[1] = {128};
        const cuuint32_t box[2] = {32, 1};                   // gather4: four boxes of one row each
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
        Pattern p{n_lines, 0, H, W, planes};
        sweep<false>(table, p, mb, tmap, !quick && items == 1.0);
        sweep<true>(table, p, mb, tmap, !quick && items == 1.0);
        CK(cudaFree(table));
    }
    return 0;
}
