// Warp-specialised fused gather + decode on the tensor cores: the production field kernel.
//
// One persistent CTA per SM, 16 warps with fixed roles, connected by mbarriers:
//
//   8 gather warps     tri-plane bilinear gather (L2/L1-bound: 24 texel lines of 128 B per sample) of a
//                      128-sample tile -> bf16 hi/lo feature tile in a 2-stage shared-memory ring
//   1 MMA warp         one elected thread issues tcgen05.mma for layer 1 (both nets) and layer 2,
//                      accumulators in a double-buffered TMEM region; tcgen05.commit signals the
//                      ring slot free and the accumulators full
//   4 epilogue warps   thread m owns TMEM lane m = sample m: tcgen05.ld, softplus, bf16 split,
//                      hidden tile -> shared (A operand of layer 2), then bias + sigmoid + stores
//
// While the epilogue warps finish tile i, the gather warps are already two tiles ahead and the
// tensor core has run layer 1 of tile i+1, so the three resources (LSU/L2, tensor pipe, ALU/SFU)
// overlap instead of taking turns as in the single-role kernel (nfe_field_tc.cu).
#include "nfe_field_launch.cuh"
#include "nfe_mlp_tc.cuh"

namespace nfe {

using namespace tcmlp;

#ifndef NFE_GATHER_WARPS
#define NFE_GATHER_WARPS 8
#endif
#ifndef NFE_PASS_CONTIG
#define NFE_PASS_CONTIG 1
#endif
constexpr int GATHER_WARPS = NFE_GATHER_WARPS;              // 8 measured best (4: 0.57, 6: 0.50, 8: 0.455, 11: 0.47 ms per pass at c2)
constexpr int EPI_WARPS = 4;
constexpr int MMA_WARP = EPI_WARPS;                         // warp index of the MMA issuer
constexpr int PIPE_THREADS = (EPI_WARPS + 1 + GATHER_WARPS) * 32;
constexpr int PASSES_PER_TILE = TILE_M / 4;                 // a pass = 4 samples (8 lanes each); passes are dealt round-robin to the gather warps
constexpr int TMEM_BUF_COLS = 192;                          // D1A 64 | D1B 64 | D2A <=48 | D2B <=32 (OSG: 48+16)
constexpr int PIPE_TMEM_COLS = 512;
// NFE_TMEM_A: the hidden activations (A operand of layer 2) go to tensor memory with tcgen05.st instead of shared memory:
// columns 384.. hold [H_A hi | H_A lo | H_B hi | H_B lo], 32 columns each (64 bf16 per row).  No 32 STS.128 + proxy fence per
// row and tile, and net B's hidden tile no longer waits for net A's layer-2 MMA to release a shared buffer.
#ifndef NFE_EPI_HALVES
#define NFE_EPI_HALVES 1  // hidden activations leave for tensor memory in halves of 32 units: 128 -> 117 registers, coarse pass 0.3788 -> 0.3764 ms at c2
#endif
#ifndef NFE_TMEM_A
#define NFE_TMEM_A 1      // measured at c2: 0.389 -> 0.381 ms per pass, and 32 KB of shared memory less
#endif
constexpr int COL_HID = 2 * TMEM_BUF_COLS;                  // 384
// Record staging: thread m parks its 192-byte record in shared memory and the warp writes its 32 records (6 KB,
// contiguous in the plain sample order) back out with fully coalesced 16-byte stores.  A direct STG.128 of 32 records
// touches 32 different lines per instruction (32 L1 wavefronts each; the record stores were a quarter of the kernel's
// L1 data-pipe load, as much as all the texel gathers); staged it is 4 + 4 + 4 (STS, LDS, STG).
// Row stride 208 B: a quarter-warp's 16-byte stores fall into 8 different bank groups.
#ifndef NFE_REC_STAGE
#define NFE_REC_STAGE 1
#endif
constexpr int REC_STAGE_STRIDE = 208;

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

template <int KIND, bool SPLIT>
struct PipeSmem {
    using T = TcTraits<KIND>;
    static constexpr int PARTS = SPLIT ? 2 : 1;
    static constexpr int NETS = T::HAS_B ? 2 : 1;
    alignas(128) unsigned char a1[2][T::SETS][PARTS][A1_BYTES];   // feature ring
    alignas(128) unsigned char b1[NETS][PARTS][B1_BYTES];
    alignas(128) unsigned char a2[PARTS][NFE_TMEM_A ? 128 : A2_BYTES];   // hidden tile (shared-memory variant), used by net A then net B
    alignas(128) unsigned char b2a[PARTS][(T::N_A / 8) * B2_SBO];
    alignas(128) unsigned char b2b[PARTS][(T::N_B / 8) * B2_SBO];
    // per sample: 12 offsets, 12 weights, item; one tile ahead.  The weights used to be parked as (w,w) pairs, FFMA2 operands
    // as loaded: 3 more LDS.128 per pass — shared-memory wavefronts are what bounds this kernel (0.422 -> 0.389 ms per pass)
    alignas(16) uint4 taps[2][GATHER_WARPS][((PASSES_PER_TILE + GATHER_WARPS - 1) / GATHER_WARPS) * 4][7];
    alignas(16) unsigned char recbuf[NFE_REC_STAGE ? TILE_M : 1][REC_STAGE_STRIDE];   // record staging (each warp owns its 32 rows)
    float bias1[NETS][HIDDEN];     // pre-multiplied by log2(e)
    float bias2a[T::N_A];
    float bias2b[T::N_B];
    alignas(8) uint64_t full[2], empty[2], d1_full[2], d2a_full[2], d2b_full[2], tmem_free[2], a2_full;
    uint32_t tmem_base;
};

// softplus in base-2 units: with t = x*log2(e), softplus(x) = ln2 * (max(t,0) + log2(1 + 2^-|t|)).
// The ln2 factor is folded into the layer-2 weights, log2(e) into the bias (one FFMA makes t).
__device__ __forceinline__ float softplus_log2(float t)
{
#ifdef NFE_ABLATE_SOFTPLUS
    return t;      // timing experiment only (wrong results): how much of the kernel is the epilogue's SFU chain
#endif
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(t)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
    return fmaxf(t, 0.0f) + l;
}

template <int KIND, bool SPLIT>
__device__ void pipe_load_params(PipeSmem<KIND, SPLIT>& s, const nfe_mlp& net_a, const nfe_mlp& net_b)
{
    using T = TcTraits<KIND>;
    constexpr int PARTS = SPLIT ? 2 : 1;
    load_weights<PARTS>(s.b1[0][0], B1_BYTES, net_a.w1, net_a.wgain1, HIDDEN, HIDDEN, FEAT, B1_LBO, B1_SBO);
    // layer-2 weights carry the ln2 of softplus_log2: fold it into the gain
    load_weights<PARTS>(s.b2a[0], sizeof(s.b2a[0]), net_a.w2, net_a.wgain2 * LN2, T::OUT_A, T::N_A, HIDDEN, B2_LBO, B2_SBO);
    for (int i = threadIdx.x; i < HIDDEN; i += blockDim.x) s.bias1[0][i] = folded_bias(net_a.b1, net_a.bgain1, i) * LOG2E;
    for (int i = threadIdx.x; i < T::N_A; i += blockDim.x) s.bias2a[i] = i < T::OUT_A ? folded_bias(net_a.b2, net_a.bgain2, i) : 0.0f;
    if constexpr (T::HAS_B) {
        load_weights<PARTS>(s.b1[1][0], B1_BYTES, net_b.w1, net_b.wgain1, HIDDEN, HIDDEN, FEAT, B1_LBO, B1_SBO);
        load_weights<PARTS>(s.b2b[0], sizeof(s.b2b[0]), net_b.w2, net_b.wgain2 * LN2, T::OUT_B, T::N_B, HIDDEN, B2_LBO, B2_SBO);
        for (int i = threadIdx.x; i < HIDDEN; i += blockDim.x) s.bias1[1][i] = folded_bias(net_b.b1, net_b.bgain1, i) * LOG2E;
        for (int i = threadIdx.x; i < T::N_B; i += blockDim.x) s.bias2b[i] = i < T::OUT_B ? folded_bias(net_b.b2, net_b.bgain2, i) : 0.0f;
    }
}

// sigmoid(x) * 1.002 - 0.001 (triplane.py:188,219,269) for two colours at once
__device__ __forceinline__ float2 rgb_activation2(float2 x)
{
    const float2 n = fmul2(x, make_float2(-LOG2E, -LOG2E));
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(n.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(n.y));
    const float2 d = fadd2(make_float2(e0, e1), make_float2(1.0f, 1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.y));
    return ffma2(make_float2(r0, r1), make_float2(1.002f, 1.002f), make_float2(-0.001f, -0.001f));
}

// hidden = softplus(D1 + b1) of one net for this thread's row, as packed bf16 parts in registers.  Pairs of
// hidden units go through packed fp32 instructions (the epilogue warps are issue-bound).
template <bool SPLIT>
__device__ __forceinline__ void hidden_to_regs(uint32_t taddr_row, const float* bias1_log2, uint32_t (&hi)[32], uint32_t (&lo)[32])
{
    const float2 k2 = make_float2(LOG2E, LOG2E), one2 = make_float2(1.0f, 1.0f), neg2 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int q = 0; q < HIDDEN / 16; ++q) {
        float v[16];
        tc::tmem_ld16(taddr_row + q * 16, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 t = ffma2(make_float2(v[2 * i], v[2 * i + 1]), k2, *reinterpret_cast<const float2*>(bias1_log2 + q * 16 + 2 * i));
            float e0, e1, l0, l1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-fabsf(t.x)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-fabsf(t.y)));
            const float2 s1 = fadd2(make_float2(e0, e1), one2);
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(s1.x));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(s1.y));
            const float2 h = fadd2(make_float2(fmaxf(t.x, 0.0f), fmaxf(t.y, 0.0f)), make_float2(l0, l1));
            const __nv_bfloat162 p = __floats2bfloat162_rn(h.x, h.y);
            hi[q * 8 + i] = *reinterpret_cast<const uint32_t*>(&p);
            if (SPLIT) {
                const float2 d = ffma2(__bfloat1622float2(p), neg2, h);      // h - bf16(h), exact
                const __nv_bfloat162 r = __floats2bfloat162_rn(d.x, d.y);
                lo[q * 8 + i] = *reinterpret_cast<const uint32_t*>(&r);
            }
        }
    }
}

// hidden_to_regs + hidden_regs_to_tmem in halves of 32 hidden units: each half's packed parts leave for tensor memory as soon
// as they exist, so 16 + 16 registers are live instead of 32 + 32 (NFE_EPI_HALVES)
template <bool SPLIT>
__device__ __forceinline__ void hidden_to_tmem_halves(uint32_t taddr_row, const float* bias1_log2, uint32_t hid_addr)
{
    const float2 k2 = make_float2(LOG2E, LOG2E), one2 = make_float2(1.0f, 1.0f), neg2 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float v[16];
            tc::tmem_ld16(taddr_row + (h * 2 + q) * 16, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 t = ffma2(make_float2(v[2 * i], v[2 * i + 1]), k2, *reinterpret_cast<const float2*>(bias1_log2 + (h * 2 + q) * 16 + 2 * i));
                float e0, e1, l0, l1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-fabsf(t.x)));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-fabsf(t.y)));
                const float2 s1 = fadd2(make_float2(e0, e1), one2);
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(s1.x));
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(s1.y));
                const float2 hh = fadd2(make_float2(fmaxf(t.x, 0.0f), fmaxf(t.y, 0.0f)), make_float2(l0, l1));
                const __nv_bfloat162 p = __floats2bfloat162_rn(hh.x, hh.y);
                hi[q * 8 + i] = *reinterpret_cast<const uint32_t*>(&p);
                if (SPLIT) {
                    const float2 d = ffma2(__bfloat1622float2(p), neg2, hh);      // h - bf16(h), exact
                    const __nv_bfloat162 r = __floats2bfloat162_rn(d.x, d.y);
                    lo[q * 8 + i] = *reinterpret_cast<const uint32_t*>(&r);
                }
            }
        }
        tc::tmem_st16(hid_addr + h * 16, hi);
        if (SPLIT) tc::tmem_st16(hid_addr + 32 + h * 16, lo);
    }
    tc::tmem_st_wait();
}

template <bool SPLIT>
__device__ __forceinline__ void hidden_regs_to_smem(unsigned char (*a2)[A2_BYTES], int row, const uint32_t (&hi)[32], const uint32_t (&lo)[32])
{
#pragma unroll
    for (int c8 = 0; c8 < HIDDEN / 8; ++c8) {
        const uint32_t off = core_offset(row, c8 * 8, A2_LBO, A2_SBO);
        *reinterpret_cast<uint4*>(a2[0] + off) = make_uint4(hi[4 * c8], hi[4 * c8 + 1], hi[4 * c8 + 2], hi[4 * c8 + 3]);
        if (SPLIT) *reinterpret_cast<uint4*>(a2[1] + off) = make_uint4(lo[4 * c8], lo[4 * c8 + 1], lo[4 * c8 + 2], lo[4 * c8 + 3]);
    }
}

// One thread issues D (+)= A * B^T with A in tensor memory (K/2 columns from tmem_a_*), B in shared memory, three split terms
template <bool SPLIT>
__device__ __forceinline__ void issue_gemm_ts(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo, const unsigned char* b_hi, const unsigned char* b_lo,
                                              int b_lbo, int b_sbo, int K, uint32_t idesc)
{
    bool acc = false;
    constexpr int TERMS = SPLIT ? 3 : 1;
#pragma unroll
    for (int t = 0; t < TERMS; ++t) {
        const uint32_t ta = (t == 1) ? tmem_a_lo : tmem_a_hi;       // hi*hi, lo*hi, hi*lo
        const unsigned char* b = (t == 2) ? b_lo : b_hi;
        for (int k = 0; k < K; k += 16) {
            const uint64_t db = tc::make_desc(tc::smem_u32(b) + (k >> 3) * b_lbo, b_lbo, b_sbo);
            tc::mma_bf16_ts(tmem_d, ta + (k >> 1), db, idesc, acc);
            acc = true;
        }
    }
}

// hidden tile of one net -> tensor memory: 32 columns of hi parts at taddr, 32 of lo parts at taddr + 32
template <bool SPLIT>
__device__ __forceinline__ void hidden_regs_to_tmem(uint32_t taddr, const uint32_t (&hi)[32], const uint32_t (&lo)[32])
{
    tc::tmem_st16(taddr, hi);
    tc::tmem_st16(taddr + 16, hi + 16);
    if (SPLIT) {
        tc::tmem_st16(taddr + 32, lo);
        tc::tmem_st16(taddr + 48, lo + 16);
    }
    tc::tmem_st_wait();
}

// Tile row -> sample.  Plain order: sample L is index L.  Quad order (FieldArgs::quad_stride): L enumerates
// (quad of 4 vertically adjacent rays) x (depth index s) x (ray in quad); the sample's storage index stays
// ray*S + s, so every consumer of the outputs is unaffected.  `item` is the batch item (plane set) of the sample.
struct SampleRef { int64_t idx; int item; };

__device__ __forceinline__ SampleRef sample_of(const FieldArgs& a, int64_t L)
{
    SampleRef r;
    if (a.quad_stride == 0) {
        r.idx = L;
        r.item = (int)(L / a.m);
        return r;
    }
    const uint32_t S = (uint32_t)a.s_per_ray, per_quad = 4u * S, res = (uint32_t)a.quad_stride;
    const uint32_t quad = (uint32_t)(L / per_quad), within = (uint32_t)(L % per_quad);
    const uint32_t s = within >> 2, ray_in_quad = within & 3u;
    const uint32_t quads_per_item = (uint32_t)(a.rays_per_item >> 2);
    const uint32_t item = quad / quads_per_item, q = quad % quads_per_item;
    const uint32_t qrow = q / res, col = q % res;
    const int64_t ray = (int64_t)item * a.rays_per_item + (int64_t)(4u * qrow + ray_in_quad) * res + col;
    r.idx = ray * S + s;
    r.item = (int)item;
    return r;
}

#ifdef NFE_PIPE_PROFILE
// Debug build only: cycles each role spends blocked on each barrier (slot k) and in total (slot 15), summed over CTAs.
__device__ unsigned long long g_pipe_prof[3][16];
#define PIPE_WAIT(slot, bar, par) do { const long long t0_ = clock64(); tc::mbar_wait(bar, par); prof_[slot] += clock64() - t0_; } while (0)
#else
#define PIPE_WAIT(slot, bar, par) tc::mbar_wait(bar, par)
#endif

// texel load at lane_base + 16*off4: one IMAD.WIDE.U32 + LDG (pointer arithmetic on a per-lane base costs four
// 64-bit ALU instructions per load otherwise)
// `streaming` taps (planes 1 and 2: no reuse between the samples a warp walks) can bypass L1 allocation
// (NFE_TAP_CG: 0 = none, 1 = ld.global.cg for those planes [default], 2 = for all taps); plane 0 keeps the read-only L1
// path, where 4 of the 12 taps hit.  Measured at c2: 0.429 -> 0.422 ms per pass (1), 0.423 (2).
#ifndef NFE_TAP_CG
#define NFE_TAP_CG 1
#endif
__device__ __forceinline__ float4 ldg_tap(const float4* lane_base, uint32_t off4, bool streaming = false)
{
#ifdef NFE_ABLATE_GATHER
    return make_float4(__uint_as_float(off4), 0.f, 0.f, 0.f);      // timing experiment only (wrong results): no texel loads
#endif
    uint64_t addr;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(off4), "l"(lane_base));
    if (NFE_TAP_CG == 2 || (NFE_TAP_CG == 1 && streaming)) return __ldcg(reinterpret_cast<const float4*>(addr));
    return __ldg(reinterpret_cast<const float4*>(addr));
}

template <int KIND, bool SPLIT>
__global__ void __launch_bounds__(PIPE_THREADS, 1) field_pipe_kernel(FieldArgs a, nfe_mlp net_a, nfe_mlp net_b)
{
    using T = TcTraits<KIND>;
    constexpr int P = SPLIT ? 1 : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeSmem<KIND, SPLIT>& s = *reinterpret_cast<PipeSmem<KIND, SPLIT>*>(smem_raw);
    // the shuffle tells the compiler the role index is warp-uniform (uniform branches / registers inside the roles)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    // ---- setup
    if (warp == MMA_WARP) tc::tmem_alloc(&s.tmem_base, PIPE_TMEM_COLS);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s.full[i], GATHER_WARPS);
            tc::mbar_init(&s.empty[i], 1);
            tc::mbar_init(&s.d1_full[i], 1);
            tc::mbar_init(&s.d2a_full[i], 1);
            tc::mbar_init(&s.d2b_full[i], 1);
            tc::mbar_init(&s.tmem_free[i], EPI_WARPS);
        }
        tc::mbar_init(&s.a2_full, EPI_WARPS);
        tc::mbar_fence_init();
    }
    pipe_load_params(s, net_a, net_b);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    const int64_t n_tiles = (a.total + TILE_M - 1) / TILE_M;
    // sigma_only with the disentangled decoder: sigma is output 0 of geo_net on the normalised planes, so the
    // de-normalised gather and the whole appearance net are skipped (gather + MLP work drops by ~55 %)
    const bool skip_b = a.sigma_only && KIND == NFE_DEC_DISENTANGLED;
#ifdef NFE_PIPE_PROFILE
    long long prof_[16] = {};
    const long long prof_t0_ = clock64();
#endif

    if (warp > MMA_WARP) {
        // ================================================================ gather warps (producers)
        // Each warp owns PER consecutive passes of the tile (a pass = 4 samples x 8 lanes); consecutive samples of
        // one ray share their (x,y)-plane texels, so they should meet in the same warp back to back.
        const int gw = warp - MMA_WARP - 1;
        const int g = lane >> 3, c4 = lane & 7;
        constexpr int PER = (PASSES_PER_TILE + GATHER_WARPS - 1) / GATHER_WARPS;
        const int n_pass = min(PER, PASSES_PER_TILE - gw * PER);             // passes this warp owns in every tile (may be <= 0)
        const int64_t set_stride4 = (int64_t)3 * a.H * a.W * (FEAT / 4);     // float4 units
        const float4* set_a = reinterpret_cast<const float4*>(a.set_norm) + c4;
        const float4* set_b = reinterpret_cast<const float4*>(a.set_denorm) + c4;
        // one plane set is read per pass in the single-gather, density-only and one-set decoders: those run the
        // rolling pipeline below; the two-set gather (24 texels per sample) has no registers left for it
        const bool affine = T::SETS == 2 && a.affine_scale != nullptr;
        const bool rolling = T::SETS == 1 || affine || skip_b;
        const float4* set_r = T::SETS == 2 ? set_a : set_b;

        // ---- tap pre-pass: ONE lane per sample computes position -> 12 clamped texel offsets + 12 weights and parks
        //      them in shared memory (the 8 lanes of a sample used to recompute them: 8x redundant issue).  It runs
        //      one tile AHEAD (double-buffered), so its dependent global loads overlap the texel loads in flight.
        auto prepass = [&](int64_t tile, int buf) {
            if (lane < PER * 4) {
                const int64_t base = tile * TILE_M;
                const int row = gw * PER * 4 + lane;
                TapSet ts;
                int item_idx = 0;
#pragma unroll
                for (int i = 0; i < 12; ++i) { ts.off4[i] = 0; ts.w[i] = 0.0f; }
                if (row < TILE_M && base + row < a.total) {
                    const SampleRef sr = sample_of(a, base + row);
                    float x, y, z;
                    if (a.coords) {
                        const float* c = a.coords + sr.idx * 3;
                        x = __ldg(c); y = __ldg(c + 1); z = __ldg(c + 2);
                    } else {
                        const int64_t ray = sr.idx / a.s_per_ray;
                        const float t = __ldg(a.depths + sr.idx);
                        const float* o = a.origins + ray * 3;
                        const float* d = a.dirs + ray * 3;
                        x = ray_point(__ldg(o), t, __ldg(d)); y = ray_point(__ldg(o + 1), t, __ldg(d + 1)); z = ray_point(__ldg(o + 2), t, __ldg(d + 2));
                    }
                    ts = make_tapset(taps3(__fmul_rn(a.scale, x), __fmul_rn(a.scale, y), __fmul_rn(a.scale, z), a.H, a.W), a.H, a.W);
                    item_idx = sr.item;
                    const int item_off = a.plane_batch == 1 ? 0 : (int)(sr.item * set_stride4);
#pragma unroll
                    for (int i = 0; i < 12; ++i) ts.off4[i] += item_off;
                }
                uint4* dst = s.taps[buf][gw][lane];
                dst[6] = make_uint4((uint32_t)item_idx, 0u, 0u, 0u);
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    dst[q] = make_uint4((uint32_t)ts.off4[4 * q], (uint32_t)ts.off4[4 * q + 1], (uint32_t)ts.off4[4 * q + 2], (uint32_t)ts.off4[4 * q + 3]);
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    dst[3 + q] = make_uint4(__float_as_uint(ts.w[4 * q]), __float_as_uint(ts.w[4 * q + 1]), __float_as_uint(ts.w[4 * q + 2]), __float_as_uint(ts.w[4 * q + 3]));
            }
        };
        auto read_taps = [&](int buf, int p, TapSet& ts) {
            const uint4* src = s.taps[buf][gw][p * 4 + g];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint4 o4 = src[q];
                ts.off4[4 * q] = (int)o4.x; ts.off4[4 * q + 1] = (int)o4.y; ts.off4[4 * q + 2] = (int)o4.z; ts.off4[4 * q + 3] = (int)o4.w;
                const uint4 w4 = src[3 + q];
                ts.w[4 * q] = __uint_as_float(w4.x); ts.w[4 * q + 1] = __uint_as_float(w4.y);
                ts.w[4 * q + 2] = __uint_as_float(w4.z); ts.w[4 * q + 3] = __uint_as_float(w4.w);
            }
        };

        float4 va[12];                    // rolling pipeline: the texels of the NEXT pass, in flight while this one is blended
        if (blockIdx.x < n_tiles && n_pass > 0) {
            prepass(blockIdx.x, 0);
            __syncwarp();
            if (rolling) {
                const uint4* src = s.taps[0][gw][g];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const uint4 o4 = src[q];
                    va[4 * q] = ldg_tap(set_r, o4.x, q != 0); va[4 * q + 1] = ldg_tap(set_r, o4.y, q != 0);
                    va[4 * q + 2] = ldg_tap(set_r, o4.z, q != 0); va[4 * q + 3] = ldg_tap(set_r, o4.w, q != 0);
                }
            }
        }
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int st = it & 1;
            const bool has_next = tile + gridDim.x < n_tiles;
            if (n_pass > 0) {
                if (has_next) prepass(tile + gridDim.x, st ^ 1);
                __syncwarp();
            }
            PIPE_WAIT(0, &s.empty[st], ((it >> 1) & 1) ^ 1);      // slot released by the layer-1 commit two tiles ago
            if (rolling) {
                for (int p = 0; p < n_pass; ++p) {
                    const int row = 4 * (gw * PER + p) + g;
                    // weights of this pass; offsets of the next one (this tile's next pass, else the next tile's first)
                    const uint4* cur = s.taps[st][gw][p * 4 + g];
                    const bool more = p + 1 < n_pass;
                    const bool fetch = more || has_next;
                    const uint4* nxt = more ? s.taps[st][gw][(p + 1) * 4 + g] : s.taps[st ^ 1][gw][g];
                    // packed fp32 pairs: channels (x,y) and (z,w) of the lane's float4; the tap weights sit in shared
                    // memory duplicated as (w,w) so they are FFMA2 operands as loaded
                    float2 f01[3], f23[3], w_in[3];
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&cur[3 + pl]);
                        const float2 w0 = make_float2(w4.x, w4.x), w1 = make_float2(w4.y, w4.y), w2 = make_float2(w4.z, w4.z), w3 = make_float2(w4.w, w4.w);
                        float2 a01 = fmul2(make_float2(va[4 * pl].x, va[4 * pl].y), w0), a23 = fmul2(make_float2(va[4 * pl].z, va[4 * pl].w), w0);
                        a01 = ffma2(make_float2(va[4 * pl + 1].x, va[4 * pl + 1].y), w1, a01); a23 = ffma2(make_float2(va[4 * pl + 1].z, va[4 * pl + 1].w), w1, a23);
                        a01 = ffma2(make_float2(va[4 * pl + 2].x, va[4 * pl + 2].y), w2, a01); a23 = ffma2(make_float2(va[4 * pl + 2].z, va[4 * pl + 2].w), w2, a23);
                        a01 = ffma2(make_float2(va[4 * pl + 3].x, va[4 * pl + 3].y), w3, a01); a23 = ffma2(make_float2(va[4 * pl + 3].z, va[4 * pl + 3].w), w3, a23);
                        f01[pl] = a01; f23[pl] = a23;
                        w_in[pl] = fadd2(fadd2(fadd2(w0, w1), w2), w3);
                        if (fetch) {             // refill the four registers just consumed with the next pass's texels
                            const uint4 o4 = nxt[pl];
                            va[4 * pl] = ldg_tap(set_r, o4.x, pl != 0); va[4 * pl + 1] = ldg_tap(set_r, o4.y, pl != 0);
                            va[4 * pl + 2] = ldg_tap(set_r, o4.z, pl != 0); va[4 * pl + 3] = ldg_tap(set_r, o4.w, pl != 0);
                        }
                    }
                    const float2 third2 = make_float2(1.0f / 3.0f, 1.0f / 3.0f);
                    const float2 fa01 = fmul2(fadd2(fadd2(f01[0], f01[1]), f01[2]), third2), fa23 = fmul2(fadd2(fadd2(f23[0], f23[1]), f23[2]), third2);
                    store_features4<SPLIT>(s.a1[st][0], row, 4 * c4, make_float4(fa01.x, fa01.y, fa23.x, fa23.y));
                    if (affine && !skip_b) {
                        // single-gather identity: only the normalised planes are read; the de-normalised features are
                        // s*f_p + m*w_in per plane (statistics: 6 L1-resident float4 loads per lane)
                        const int item = a.affine_items == 1 ? 0 : (int)cur[6].x;
                        const float4* sc = reinterpret_cast<const float4*>(a.affine_scale + (int64_t)item * 96) + c4;
                        const float4* sh = reinterpret_cast<const float4*>(a.affine_shift + (int64_t)item * 96) + c4;
                        float2 d01[3], d23[3];
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl) {
                            const float4 scl = __ldg(sc + pl * 8), shf = __ldg(sh + pl * 8);
                            d01[pl] = ffma2(make_float2(scl.x, scl.y), f01[pl], fmul2(make_float2(shf.x, shf.y), w_in[pl]));
                            d23[pl] = ffma2(make_float2(scl.z, scl.w), f23[pl], fmul2(make_float2(shf.z, shf.w), w_in[pl]));
                        }
                        const float2 fb01 = fmul2(fadd2(fadd2(d01[0], d01[1]), d01[2]), third2), fb23 = fmul2(fadd2(fadd2(d23[0], d23[1]), d23[2]), third2);
                        store_features4<SPLIT>(s.a1[st][T::SETS - 1], row, 4 * c4, make_float4(fb01.x, fb01.y, fb23.x, fb23.y));
                    }
                }
            } else {
                for (int p = 0; p < n_pass; ++p) {
                    const int row = 4 * (gw * PER + p) + g;
                    TapSet ts;
                    read_taps(st, p, ts);
                    // all 24 texel loads of the sample (two plane sets) are issued before the first blend
                    float4 vb[12];
#pragma unroll
                    for (int i = 0; i < 12; ++i) va[i] = ldg_tap(set_a, (uint32_t)ts.off4[i]);
#pragma unroll
                    for (int i = 0; i < 12; ++i) vb[i] = ldg_tap(set_b, (uint32_t)ts.off4[i]);
                    store_features4<SPLIT>(s.a1[st][0], row, 4 * c4, gather_reduce(va, ts));
                    store_features4<SPLIT>(s.a1[st][T::SETS - 1], row, 4 * c4, gather_reduce(vb, ts));
                }
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.full[st]);
        }
    } else if (warp == MMA_WARP) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc1 = tc::make_idesc_bf16(TILE_M, HIDDEN);
            constexpr uint32_t idesc2a = tc::make_idesc_bf16(TILE_M, T::N_A);
            constexpr uint32_t idesc2b = tc::make_idesc_bf16(TILE_M, T::N_B);
            // layer 1 of tile `it` into ring slot / TMEM buffer it&1
            auto layer1 = [&](int it) {
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                const uint32_t tb = tmem + st * TMEM_BUF_COLS;
                PIPE_WAIT(1, &s.tmem_free[st], ph ^ 1);            // epilogue is done with this TMEM buffer (tile it-2)
                PIPE_WAIT(2, &s.full[st], ph);                     // features landed
                tc::fence_after_sync();
                issue_gemm<SPLIT>(tb + COL_D1A, s.a1[st][0][0], s.a1[st][0][P], A1_LBO, A1_SBO, s.b1[0][0], s.b1[0][P], B1_LBO, B1_SBO, FEAT, idesc1);
                if (T::HAS_B && !skip_b)
                    issue_gemm<SPLIT>(tb + COL_D1B, s.a1[st][T::SETS - 1][0], s.a1[st][T::SETS - 1][P], A1_LBO, A1_SBO, s.b1[1][0], s.b1[1][P],
                                      B1_LBO, B1_SBO, FEAT, idesc1);
                tc::mma_commit(&s.empty[st]);                       // ring slot reusable once these MMAs have read it
                tc::mma_commit(&s.d1_full[st]);
            };
            int it = 0;
            uint32_t a2_uses = 0;
            if ((int64_t)blockIdx.x < n_tiles) layer1(0);
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t tb = tmem + (it & 1) * TMEM_BUF_COLS;
                // layer 2, net A
                PIPE_WAIT(3, &s.a2_full, a2_uses++ & 1);
                tc::fence_after_sync();
#if NFE_TMEM_A
                issue_gemm_ts<SPLIT>(tb + COL_D2A, tmem + COL_HID, tmem + COL_HID + 32, s.b2a[0], s.b2a[P], B2_LBO, B2_SBO, HIDDEN, idesc2a);
#else
                issue_gemm<SPLIT>(tb + COL_D2A, s.a2[0], s.a2[P], A2_LBO, A2_SBO, s.b2a[0], s.b2a[P], B2_LBO, B2_SBO, HIDDEN, idesc2a);
#endif
                tc::mma_commit(&s.d2a_full[it & 1]);
                // layer 1 of the NEXT tile goes in here, so the epilogue never waits on it
                if (tile + gridDim.x < n_tiles) layer1(it + 1);
                if (T::HAS_B && !skip_b) {
                    PIPE_WAIT(3, &s.a2_full, a2_uses++ & 1);
                    tc::fence_after_sync();
#if NFE_TMEM_A
                    issue_gemm_ts<SPLIT>(tb + COL_D2A + T::N_A, tmem + COL_HID + 64, tmem + COL_HID + 96, s.b2b[0], s.b2b[P], B2_LBO, B2_SBO, HIDDEN, idesc2b);
#else
                    issue_gemm<SPLIT>(tb + COL_D2A + T::N_A, s.a2[0], s.a2[P], A2_LBO, A2_SBO, s.b2b[0], s.b2b[P], B2_LBO, B2_SBO, HIDDEN, idesc2b);
#endif
                    tc::mma_commit(&s.d2b_full[it & 1]);
                }
            }
        }
        __syncwarp();
    } else {
        // ================================================================ epilogue warps (TMEM lanes 32*warp ..)
        const int row = threadIdx.x;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int st = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const uint32_t lane_addr = tmem + st * TMEM_BUF_COLS + ((uint32_t)(warp * 32) << 16);
            uint32_t hi[32], lo[32];
            PIPE_WAIT(4, &s.d1_full[st], ph);
            tc::fence_after_sync();
#if NFE_TMEM_A && NFE_EPI_HALVES
            // the previous tile's last layer-2 MMA has completed (we waited on its commit), so the hidden tile is free
            const uint32_t hid_addr = tmem + COL_HID + ((uint32_t)(warp * 32) << 16);
            hidden_to_tmem_halves<SPLIT>(lane_addr + COL_D1A, s.bias1[0], hid_addr);
            tc::fence_before_sync();
#elif NFE_TMEM_A
            hidden_to_regs<SPLIT>(lane_addr + COL_D1A, s.bias1[0], hi, lo);
            // the previous tile's last layer-2 MMA has completed (we waited on its commit), so the hidden tile is free
            const uint32_t hid_addr = tmem + COL_HID + ((uint32_t)(warp * 32) << 16);
            hidden_regs_to_tmem<SPLIT>(hid_addr, hi, lo);
            tc::fence_before_sync();
#else
            hidden_to_regs<SPLIT>(lane_addr + COL_D1A, s.bias1[0], hi, lo);
            hidden_regs_to_smem<SPLIT>(s.a2, row, hi, lo);
            tc::fence_async_smem();
#endif
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.a2_full);
            if (T::HAS_B && !skip_b) {
#if NFE_TMEM_A && NFE_EPI_HALVES
                hidden_to_tmem_halves<SPLIT>(lane_addr + COL_D1B, s.bias1[1], hid_addr + 64);   // overlaps the net-A layer-2 MMA; its own columns
#else
                hidden_to_regs<SPLIT>(lane_addr + COL_D1B, s.bias1[1], hi, lo);     // overlaps the net-A layer-2 MMA
#endif
#if NFE_TMEM_A && NFE_EPI_HALVES
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s.a2_full);
                PIPE_WAIT(5, &s.d2a_full[st], ph);
                tc::fence_after_sync();
#elif NFE_TMEM_A
                hidden_regs_to_tmem<SPLIT>(hid_addr + 64, hi, lo);                // its own columns: no wait for net A's MMA
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s.a2_full);
                PIPE_WAIT(5, &s.d2a_full[st], ph);
                tc::fence_after_sync();
#else
                PIPE_WAIT(5, &s.d2a_full[st], ph);                                // net A consumed the hidden tile
                tc::fence_after_sync();
                hidden_regs_to_smem<SPLIT>(s.a2, row, hi, lo);
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s.a2_full);
#endif
            } else {
                PIPE_WAIT(5, &s.d2a_full[st], ph);
                tc::fence_after_sync();
            }
            // ---- outputs
            const bool live = tile * TILE_M + row < a.total;
            const int64_t idx = live ? sample_of(a, tile * TILE_M + row).idx : 0;
            float outa[T::N_A];
#pragma unroll
            for (int q = 0; q < T::N_A / 16; ++q) {
                float v[16];
                tc::tmem_ld16(lane_addr + COL_D2A + q * 16, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 o2 = fadd2(make_float2(v[2 * i], v[2 * i + 1]), *reinterpret_cast<const float2*>(&s.bias2a[q * 16 + 2 * i]));
                    outa[q * 16 + 2 * i] = o2.x; outa[q * 16 + 2 * i + 1] = o2.y;
                }
            }
            float sig = outa[0];
            if (a.density_noise > 0.0f && live) {
                const uint4 r = philox4x32(a.seed, (uint64_t)idx, a.offset);
                sig += normal2(r.x, r.y).x * a.density_noise;
            }
            if (live) a.sigma[idx] = sig;
            if (a.sigma_only) {                                       // nothing else is written
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s.tmem_free[st]);
                continue;
            }
            // staged records need the tile's rows to be consecutive samples (plain order)
            const bool staged = NFE_REC_STAGE && a.rec && a.quad_stride == 0;
            float4* rec = a.rec ? (staged ? reinterpret_cast<float4*>(s.recbuf[row]) : reinterpret_cast<float4*>(a.rec + idx * 48)) : nullptr;
            float4* rgb4 = rec ? rec + 4 : reinterpret_cast<float4*>(a.rgb + idx * 32);
            if constexpr (KIND == NFE_DEC_DISENTANGLED) {
                if (live) {
                    if (rec) {
                        rec[0] = make_float4(sig, outa[1], outa[2], outa[3]);
#pragma unroll
                        for (int c = 1; c < 4; ++c) rec[c] = make_float4(outa[4 * c], outa[4 * c + 1], outa[4 * c + 2], outa[4 * c + 3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 15; ++c) a.seg[idx * 15 + c] = outa[1 + c];
                    }
                }
            } else {
                if (live) {
                    if (rec && !T::HAS_B) {
                        rec[0] = make_float4(sig, 0.f, 0.f, 0.f);
                        rec[1] = rec[2] = rec[3] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        rgb4[c] = make_float4(rgb_activation(outa[1 + 4 * c]), rgb_activation(outa[2 + 4 * c]),
                                              rgb_activation(outa[3 + 4 * c]), rgb_activation(outa[4 + 4 * c]));
                }
            }
            if constexpr (T::HAS_B) {
                PIPE_WAIT(6, &s.d2b_full[st], ph);
                tc::fence_after_sync();
                float outb[T::N_B];
#pragma unroll
                for (int q = 0; q < T::N_B / 16; ++q) {
                    float v[16];
                    tc::tmem_ld16(lane_addr + COL_D2A + T::N_A + q * 16, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 o2 = fadd2(make_float2(v[2 * i], v[2 * i + 1]), *reinterpret_cast<const float2*>(&s.bias2b[q * 16 + 2 * i]));
                        outb[q * 16 + 2 * i] = o2.x; outb[q * 16 + 2 * i + 1] = o2.y;
                    }
                }
                if (live) {
                    if constexpr (KIND == NFE_DEC_DISENTANGLED) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float2 lo2 = rgb_activation2(make_float2(outb[4 * c], outb[4 * c + 1]));
                            const float2 hi2 = rgb_activation2(make_float2(outb[4 * c + 2], outb[4 * c + 3]));
                            rgb4[c] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
                        }
                    } else if (rec) {
                        rec[0] = make_float4(sig, outb[0], outb[1], outb[2]);
#pragma unroll
                        for (int c = 1; c < 4; ++c) rec[c] = make_float4(outb[4 * c - 1], outb[4 * c], outb[4 * c + 1], outb[4 * c + 2]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 15; ++c) a.seg[idx * 15 + c] = outb[c];
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.tmem_free[st]);
            if (staged) {
                // 16-byte chunk g of the warp's 32 records sits at row g/12, chunk g%12 of the staging rows
                const int64_t row0 = tile * TILE_M + warp * 32;
                const int n_chunks = (int)min((int64_t)32, a.total - row0) * 12;
                float4* gdst = reinterpret_cast<float4*>(a.rec + row0 * 48);
                int r = lane / 12, c = lane % 12;
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const int g = k * 32 + lane;
                    const float4 v = *reinterpret_cast<const float4*>(s.recbuf[warp * 32 + r] + c * 16);
                    if (g < n_chunks) gdst[g] = v;
                    r += 2; c += 8;
                    if (c >= 12) { c -= 12; r += 1; }
                }
                __syncwarp();          // the rows are rewritten by the next tile
            }
        }
    }

#ifdef NFE_PIPE_PROFILE
    if (lane == 0) {
        const int role = warp > MMA_WARP ? 0 : (warp == MMA_WARP ? 1 : 2);
        prof_[15] = clock64() - prof_t0_;
        for (int i = 0; i < 16; ++i) atomicAdd(&g_pipe_prof[role][i], (unsigned long long)prof_[i]);
    }
#endif
    // ---- teardown
    tc::fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc(tmem, PIPE_TMEM_COLS);
}

template <int KIND, bool SPLIT>
static int launch_field_pipe_kind(const FieldArgs& a, const nfe_mlp& net_a, const nfe_mlp& net_b, cudaStream_t stream)
{
    const size_t smem = sizeof(PipeSmem<KIND, SPLIT>) + 128;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(field_pipe_kernel<KIND, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("field_pipe_kernel: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return 2; }
        configured = true;
    }
    const int64_t n_tiles = (a.total + TILE_M - 1) / TILE_M;
    const int64_t cap = sm_count();   // persistent: one CTA per SM (it owns all 512 TMEM columns)
    field_pipe_kernel<KIND, SPLIT><<<(unsigned)(n_tiles < cap ? n_tiles : cap), PIPE_THREADS, smem, stream>>>(a, net_a, net_b);
    return check_launch("field_pipe_kernel");
}

int launch_field_pipe(int kind, int precision, const FieldArgs& a, const nfe_mlp* net_a, const nfe_mlp* net_b, cudaStream_t stream)
{
    if (a.total <= 0) return 0;
    nfe_mlp none = {};
    const bool split = precision == NFE_PREC_BF16X3;
    switch (kind) {
        case NFE_DEC_OSG:
            return split ? launch_field_pipe_kind<NFE_DEC_OSG, true>(a, *net_a, none, stream) : launch_field_pipe_kind<NFE_DEC_OSG, false>(a, *net_a, none, stream);
        case NFE_DEC_DISENTANGLED:
            return split ? launch_field_pipe_kind<NFE_DEC_DISENTANGLED, true>(a, *net_a, *net_b, stream)
                         : launch_field_pipe_kind<NFE_DEC_DISENTANGLED, false>(a, *net_a, *net_b, stream);
        default:
            return split ? launch_field_pipe_kind<NFE_DEC_SEGMENTATION, true>(a, *net_a, *net_b, stream)
                         : launch_field_pipe_kind<NFE_DEC_SEGMENTATION, false>(a, *net_a, *net_b, stream);
    }
}

}  // namespace nfe

#ifdef NFE_PIPE_PROFILE
NFE_EXPORT int nfe_debug_pipe_profile(unsigned long long* out48, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out48, nfe::g_pipe_prof, sizeof(unsigned long long) * 48);
    if (reset) { unsigned long long z[48] = {}; cudaMemcpyToSymbol(nfe::g_pipe_prof, z, sizeof(z)); }
    return 0;
}
#endif
