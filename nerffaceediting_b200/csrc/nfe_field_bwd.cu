// Fused backward of gather + DisentangledOSGDecoder on the tensor cores (BASELINE config 4).
//
// Per 128-sample tile, one CTA (256 threads) does what the reference's autograd graph does in ~40 kernels
// (grid_sampler backward, two addmm/softplus/addmm/sigmoid chains, triplane.py:249-270):
//
//   gather        X = [mean features of the normalised planes | of the raw planes | 1]      (128 x 80, bf16 hi/lo)
//   G1            pre_n = X_n W1_n^T                                                          (tcgen05, TMEM)
//   epilogue 1    H_n = softplus(pre_n + b1_n) -> smem;  dY_n from the per-sample record gradients
//                 (geo: d sigma, d seg; app: d rgb * 1.002 * s(1-s), s recovered from the saved rgb)
//   G2            dH_n = dY_n W2_n                              G4   dW2 += H^T dY   (operands read MN-major)
//   epilogue 2    dpre_n = dH_n * sigmoid(pre_n + b1_n) -> smem (over H)
//   G3            dX_n = dpre_n W1_n                            G5   dW1 | db1 += dpre^T [X | 1]
//   epilogue 3    dX -> smem (fp32) -> red.global.add.v4.f32 of w_tap/3 * dX into the channel-last plane gradients
//
// dW1/db1/dW2 accumulate in TMEM over all tiles of the CTA and are added to the global gradients once at the end
// (db2 in registers).  Every product runs as three bf16 MMAs (hi*hi + lo*hi + hi*lo, fp32 accumulate), like the
// forward's bf16x3 mode.  The two weight-gradient GEMMs contract over the tile's ROWS: their operands are the same
// shared-memory tiles the other GEMMs read K-major, described MN-major (transposed) to the tensor core, so nothing
// is transposed in software.
//
// KIND selects the decoder (triplane.py:167-270).  DISENTANGLED is the layout described above.  SEGMENTATION has `net`
// (33 outputs: sigma + 32 colours) and `seg_net` (15 logits), BOTH on the raw-plane features: the raw set is gathered once and
// written into both X column groups, dY = [dY_net (48, 33 used) | dY_seg (16, 15 used)], and dX_net + dX_seg scatter into the
// raw-plane gradient.  OSG is SEGMENTATION without a second net (its weights are zeros, nothing of it is flushed).
//
// AFFINE = true is the backward of the forward's single-gather identity: the raw planes are known to be
// norm*scale + shift per (item, plane, channel), so only the normalised set is gathered and only its gradient is
// scattered.  With f_p the per-plane blend of the normalised planes and w_p the in-bounds tap-weight sum,
//   X_raw = mean_p(scale_p f_p + shift_p w_p)   =>   d f_p = (dX_norm + scale_p dX_raw)/3,
//   d scale_p = sum_samples f_p dX_raw / 3,      d shift_p = sum_samples w_p dX_raw / 3
// (the two statistics gradients accumulate in registers per batch item and leave as a handful of atomics).  Gather and
// scatter traffic — the two phases that dominate the kernel — halve.
#include <cuda_fp16.h>
#include "nfe_field.cuh"
#include "nfe_mlp_tc.cuh"

namespace nfe {

using namespace tcmlp;

namespace fb {

constexpr int THREADS = 256;
constexpr int XA_COLS = 80, XA_LBO = 160, XA_SBO = (XA_COLS / 8) * XA_LBO, XA_BYTES = 16 * XA_SBO;   // [X_norm | X_raw | 1 0..0]
constexpr int HC_COLS = 128, HC_LBO = 128, HC_SBO = (HC_COLS / 8) * HC_LBO, HC_BYTES = 16 * HC_SBO;  // [H_geo | H_app], later dpre
// output-layer layout per decoder kind: real / padded output counts of the two nets; dY = [dY_0 (PAD0) | dY_1 (PAD1)]
template <int KIND> struct Out;
template <> struct Out<NFE_DEC_DISENTANGLED> { static constexpr int OUT0 = 16, PAD0 = 16, OUT1 = 32, PAD1 = 32; };   // geo | app
template <> struct Out<NFE_DEC_SEGMENTATION> { static constexpr int OUT0 = 33, PAD0 = 48, OUT1 = 15, PAD1 = 16; };   // net | seg_net
template <> struct Out<NFE_DEC_OSG> { static constexpr int OUT0 = 33, PAD0 = 48, OUT1 = 0, PAD1 = 16; };              // net | (none)
constexpr int DY_LBO = 128;
constexpr int W2T_LBO = 128;                                             // B of G2: [N=64 x K=PADn], element (j,o) = W2[o][j]
constexpr int W1T_LBO = 128, W1T_SBO = 1024, W1T_BYTES = 4 * W1T_SBO;    // B of G3: [N=32 x K=64], element (i,j) = W1[j][i]
constexpr int GX_STRIDE = 68;                                            // fp32 dX staging row stride (floats), aliases hc
// TMEM columns
constexpr int C_PRE = 0, C_DH = 128, C_DX = 256, C_DW1 = 320, C_DW2 = 400, TMEM_ALLOC = 512;

template <int KIND, bool AFFINE>
struct Smem {
    static constexpr int DY_COLS = Out<KIND>::PAD0 + Out<KIND>::PAD1, DY_SBO = (DY_COLS / 8) * DY_LBO, DY_BYTES = 16 * DY_SBO;
    static constexpr int KMAX = Out<KIND>::PAD0 > Out<KIND>::PAD1 ? Out<KIND>::PAD0 : Out<KIND>::PAD1;
    static constexpr int W2T_SBO = (KMAX / 8) * W2T_LBO, W2T_BYTES = 8 * W2T_SBO;        // 512 / 4 KB (disentangled), 768 / 6 KB (33 outputs)
    alignas(128) unsigned char xa[2][XA_BYTES];
    alignas(128) unsigned char hc[2][HC_BYTES];
    alignas(128) unsigned char dy[2][DY_BYTES];
    alignas(128) unsigned char w1[2][2][B1_BYTES];
    alignas(128) unsigned char w2t[2][2][W2T_BYTES];
    alignas(128) unsigned char w1t[2][2][W1T_BYTES];
    alignas(16) int tap_off[TILE_M][12];
    alignas(16) float tap_w[TILE_M][12];
    int tap_item[TILE_M];
    alignas(16) __half fp[AFFINE ? TILE_M : 1][96];      // per-plane blends of the normalised planes, kept for d scale
    float bias1[2][HIDDEN];
    alignas(8) uint64_t bar;
    uint32_t tmem_base;
};
static_assert(GX_STRIDE * 4 * TILE_M <= 2 * HC_BYTES, "dX staging must fit in the hidden tile it aliases");

struct Args {
    const float* set_norm; const float* set_raw;      // channel-last plane sets
    float* g_norm; float* g_raw;                      // channel-last plane gradients (zero-initialised by the caller)
    int plane_batch, H, W; float scale;
    const float* origins; const float* dirs; const float* depths;
    const float* coords;                              // explicit sample positions [total,3] (run_model); rays are unused then
    int s_per_ray; int64_t m, total;
    const float* rec;                                 // forward records [total,48] (rgb for the sigmoid derivative)
    const float* g_rec;                               // d loss / d record [total,48]
    float* gw1[2]; float* gb1[2]; float* gw2[2]; float* gb2[2];    // raw-parameter gradients (accumulated atomically)
    // AFFINE: raw = norm*scale + shift, [affine_items, 96] each (affine_items = batch or 1), and their gradients
    const float* affine_scale; const float* affine_shift; int affine_items;
    float* g_scale; float* g_shift;
};

// instruction descriptor with both operands MN-major (bits 15/16)
__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) { return tc::make_idesc_bf16(M, N) | (1u << 15) | (1u << 16); }

// D (+)= A^T-view * B^T-view contracting over the 128 tile rows; both tiles are stored K-major [rows x cols] with
// (lbo,sbo); read MN-major the roles swap: stride between 8-row K blocks = sbo, between 8-column MN blocks = lbo
__device__ __forceinline__ void issue_gemm_rows(bool leader, uint32_t tmem_d, const unsigned char* a_hi, const unsigned char* a_lo, int a_lbo, int a_sbo,
                                                const unsigned char* b_hi, const unsigned char* b_lo, int b_lbo, int b_sbo, uint32_t idesc, bool accumulate)
{
    // every lane of the issuing warp runs this (warp-uniform arguments), the elected one issues; descriptors = constant part + an add
    const uint64_t da0 = tc::make_desc(0, a_sbo, a_lbo), db0 = tc::make_desc(0, b_sbo, b_lbo);
    const uint32_t ah = tc::smem_u32(a_hi) >> 4, al = tc::smem_u32(a_lo) >> 4, bh = tc::smem_u32(b_hi) >> 4, bl = tc::smem_u32(b_lo) >> 4;
    const uint32_t ak = (uint32_t)(2 * a_sbo) >> 4, bk = (uint32_t)(2 * b_sbo) >> 4;
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const uint32_t a = (t == 1) ? al : ah, b = (t == 2) ? bl : bh;
#pragma unroll
        for (int ks = 0; ks < TILE_M / 16; ++ks) {
            if (leader) tc::mma_bf16_ss(tmem_d, da0 + (a + ks * ak), db0 + (b + ks * bk), idesc, acc);
            acc = 1u;
        }
    }
}

// transposed, gain-folded weight operand: dst element (n, k) = w[k * ld + n] * gain for n < N, k < K (zero elsewhere)
__device__ void load_weights_t(unsigned char* dst, size_t part_stride, const float* w, float gain, int N, int K, int K_pad, int ld, int lbo, int sbo)
{
    for (int i = threadIdx.x; i < N * K_pad; i += blockDim.x) {
        const int n = i / K_pad, k = i % K_pad;
        const float v = k < K ? __fmul_rn(__ldg(w + k * ld + n), gain) : 0.0f;
        __nv_bfloat16 hi, lo;
        tc::split_bf16(v, hi, lo);
        const uint32_t off = core_offset(n, k, lbo, sbo);
        *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(dst + part_stride + off) = lo;
    }
}

// 8 consecutive columns of one row as bf16 hi / lo parts (two 16-byte stores)
__device__ __forceinline__ void store8(unsigned char* hi_tile, unsigned char* lo_tile, uint32_t off, const float (&v)[8])
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 p = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float2 back = __bfloat1622float2(p);
        const __nv_bfloat162 r = __floats2bfloat162_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&p);
        l[i] = *reinterpret_cast<const uint32_t*>(&r);
    }
    *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// texel loads with the streaming planes (1 and 2: no reuse between the samples a warp walks) bypassing L1 allocation
__device__ __forceinline__ void gather_load_streaming(const float* __restrict__ set, const TapSet& ts, int c4, float4 (&v)[12])
{
    const float4* base = reinterpret_cast<const float4*>(set) + c4;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ldg(base + ts.off4[i]);
#pragma unroll
    for (int i = 4; i < 12; ++i) v[i] = __ldcg(base + ts.off4[i]);
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifdef NFE_BWD_PROFILE
// Debug build only: cycles thread 0 of every CTA spends in each phase of the tile loop, summed over CTAs.
__device__ unsigned long long g_bwd_prof[16];
#define BWD_MARK(slot) do { if (threadIdx.x == 0) { const long long t_ = clock64(); prof_[slot] += t_ - prof_t_; prof_t_ = t_; } } while (0)
#else
#define BWD_MARK(slot) do { } while (0)
#endif

template <int KIND, bool AFFINE>
__global__ void __launch_bounds__(THREADS, 1) field_bwd_kernel(Args a, nfe_mlp geo, nfe_mlp app)
{
    using O = Out<KIND>;
    constexpr bool DIS = KIND == NFE_DEC_DISENTANGLED;
    constexpr int DY_COLS = O::PAD0 + O::PAD1, DY_SBO = (DY_COLS / 8) * DY_LBO;
    constexpr int W2T_SBO = Smem<KIND, AFFINE>::W2T_SBO, W2T_BYTES = Smem<KIND, AFFINE>::W2T_BYTES;
    static_assert(!AFFINE || DIS, "the single-gather identity belongs to the disentangled decoder");
    static_assert(C_DW2 + DY_COLS <= TMEM_ALLOC, "tensor memory");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // (no integer round-up of the base: it strips the shared address space and every LDS/STS below becomes a generic LD/ST — round 1's
    // build had 258 of those and 10 LDS/STS; the dynamic window starts 1024-byte aligned, the kernel has no static shared memory)
    Smem<KIND, AFFINE>& s = *reinterpret_cast<Smem<KIND, AFFINE>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int net = warp >> 2;                       // epilogue role: warps 0-3 geo_net, 4-7 app_net
    const int row = (warp & 3) * 32 + lane;          // TMEM lane = tile row
    const nfe_mlp& mine = net ? app : geo;

    // ---- setup: TMEM, barrier, weights (three layouts), the constant [1 0 .. 0] block of X
    if (warp == 0) tc::tmem_alloc(&s.tmem_base, TMEM_ALLOC);
    if (threadIdx.x == 0) { tc::mbar_init(&s.bar, 1); tc::mbar_fence_init(); }
    for (int n = 0; n < 2; ++n) {
        const nfe_mlp& p = n ? app : geo;
        const bool present = n == 0 || O::OUT1 > 0;            // OSG has no second net: all-zero operands, nothing flushed
        load_weights<2>(s.w1[n][0], B1_BYTES, p.w1, p.wgain1, present ? HIDDEN : 0, HIDDEN, FEAT, B1_LBO, B1_SBO);
        load_weights_t(s.w2t[n][0], W2T_BYTES, p.w2, p.wgain2, HIDDEN, n ? O::OUT1 : O::OUT0, n ? O::PAD1 : O::PAD0, HIDDEN, W2T_LBO, W2T_SBO);
        load_weights_t(s.w1t[n][0], W1T_BYTES, p.w1, p.wgain1, FEAT, present ? HIDDEN : 0, HIDDEN, FEAT, W1T_LBO, W1T_SBO);
        for (int i = threadIdx.x; i < HIDDEN; i += blockDim.x) s.bias1[n][i] = present ? folded_bias(p.b1, p.bgain1, i) : 0.0f;
    }
    for (int i = threadIdx.x; i < TILE_M * 16; i += blockDim.x) {
        const int r = i >> 4, c = 64 + (i & 15);
        const uint32_t off = core_offset(r, c, XA_LBO, XA_SBO);
        *reinterpret_cast<__nv_bfloat16*>(s.xa[0] + off) = __float2bfloat16_rn(c == 64 ? 1.0f : 0.0f);
        *reinterpret_cast<__nv_bfloat16*>(s.xa[1] + off) = __float2bfloat16_rn(0.0f);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;
    constexpr int DB2 = O::PAD0 > O::PAD1 ? O::PAD0 : O::PAD1;
    float db2[DB2];
#pragma unroll
    for (int i = 0; i < DB2; ++i) db2[i] = 0.0f;

    const int64_t n_tiles = (a.total + TILE_M - 1) / TILE_M;
    const int64_t set_stride4 = (int64_t)3 * a.H * a.W * (FEAT / 4);
    const int c4 = threadIdx.x & 7;
    bool first = true;
    // AFFINE: statistics gradients of this thread's 4 channels x 3 planes for batch item `cur_item`
    float4 ds[3], dm[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) ds[p] = dm[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur_item = -1;
    auto flush_stats = [&]() {
        if (!AFFINE || cur_item < 0) return;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            float vs[8] = {ds[p].x, ds[p].y, ds[p].z, ds[p].w, dm[p].x, dm[p].y, dm[p].z, dm[p].w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {       // the 4 sample groups of the warp hold the same channels
                vs[i] += __shfl_xor_sync(0xffffffffu, vs[i], 8);
                vs[i] += __shfl_xor_sync(0xffffffffu, vs[i], 16);
            }
            if (lane < 8) {
                float* gs = a.g_scale + (int64_t)cur_item * 96 + p * 32 + 4 * c4;
                float* gm = a.g_shift + (int64_t)cur_item * 96 + p * 32 + 4 * c4;
#pragma unroll
                for (int i = 0; i < 4; ++i) { atomicAdd(gs + i, vs[i]); atomicAdd(gm + i, vs[4 + i]); }
            }
            ds[p] = dm[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    // ---- tap pre-pass: ONE thread per sample turns (ray, depth) into 12 clamped texel offsets + 12 weights (kept in shared memory
    //      for the scatter at the end of the tile).  Tile i + 1's pre-pass is COMPUTED in the shadow of tile i's first GEMM by the
    //      upper 128 threads (which only wait there) and held in their registers until tile i's scatter has read the tables: its
    //      dependent depth / ray loads used to sit at the head of every tile (shared memory has no room for a second table)
    auto tap_compute = [&](int64_t tbase, TapSet& ts, int& item) {
        const int64_t idx = tbase + (threadIdx.x & (TILE_M - 1));
#pragma unroll
        for (int i = 0; i < 12; ++i) { ts.off4[i] = 0; ts.w[i] = 0.0f; }
        item = 0;
        if (idx < a.total) {
            float x, y, z;
            if (a.coords) {
                const float* c = a.coords + idx * 3;
                x = __ldg(c); y = __ldg(c + 1); z = __ldg(c + 2);
            } else {
                const int64_t ray = idx / a.s_per_ray;
                const float t = __ldg(a.depths + idx);
                const float* o = a.origins + ray * 3;
                const float* d = a.dirs + ray * 3;
                x = ray_point(__ldg(o), t, __ldg(d)); y = ray_point(__ldg(o + 1), t, __ldg(d + 1)); z = ray_point(__ldg(o + 2), t, __ldg(d + 2));
            }
            ts = make_tapset(taps3(__fmul_rn(a.scale, x), __fmul_rn(a.scale, y), __fmul_rn(a.scale, z), a.H, a.W), a.H, a.W);
            item = (int)(idx / a.m);
            const int item_off = a.plane_batch == 1 ? 0 : (int)(item * set_stride4);
#pragma unroll
            for (int i = 0; i < 12; ++i) ts.off4[i] += item_off;
        }
    };
    auto tap_store = [&](const TapSet& ts, int item) {
        const int r = threadIdx.x & (TILE_M - 1);
        s.tap_item[r] = item;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            *reinterpret_cast<int4*>(&s.tap_off[r][4 * q]) = make_int4(ts.off4[4 * q], ts.off4[4 * q + 1], ts.off4[4 * q + 2], ts.off4[4 * q + 3]);
            *reinterpret_cast<float4*>(&s.tap_w[r][4 * q]) = make_float4(ts.w[4 * q], ts.w[4 * q + 1], ts.w[4 * q + 2], ts.w[4 * q + 3]);
        }
    };
    TapSet next_ts;
    int next_item = 0;
#ifdef NFE_BWD_PROFILE
    long long prof_[16] = {};
    long long prof_t_ = clock64();
#endif
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE_M;
        BWD_MARK(9);
        // AFFINE: which item's statistics this tile feeds (a tile straddling two items goes sample by sample)
        bool straddle = false;
        if (AFFINE) {
            const int64_t last = (base + TILE_M < a.total ? base + TILE_M : a.total) - 1;
            const int it_first = a.affine_items == 1 ? 0 : (int)(base / a.m), it_last = a.affine_items == 1 ? 0 : (int)(last / a.m);
            straddle = it_first != it_last;
            if (!straddle && it_first != cur_item) { flush_stats(); cur_item = it_first; }
        }
        if (tile == (int64_t)blockIdx.x) {                   // the first tile has nobody to hide behind
            if (threadIdx.x >= TILE_M) tap_compute(base, next_ts, next_item);
        }
        if (threadIdx.x >= TILE_M) tap_store(next_ts, next_item);
        __syncthreads();
        // ---- gather (8 lanes per sample, 32 samples per step; AFFINE: two steps' texel loads in flight together)
        auto read_taps = [&](int r, TapSet& ts) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int4 o4 = *reinterpret_cast<const int4*>(&s.tap_off[r][4 * q]);
                const float4 w4 = *reinterpret_cast<const float4*>(&s.tap_w[r][4 * q]);
                ts.off4[4 * q] = o4.x; ts.off4[4 * q + 1] = o4.y; ts.off4[4 * q + 2] = o4.z; ts.off4[4 * q + 3] = o4.w;
                ts.w[4 * q] = w4.x; ts.w[4 * q + 1] = w4.y; ts.w[4 * q + 2] = w4.z; ts.w[4 * q + 3] = w4.w;
            }
        };
        auto store_x = [&](int r, const float4& fa, const float4& fb) {
#pragma unroll
            for (int set = 0; set < 2; ++set) {
                const float4 f = set ? fb : fa;
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(f.x, f.y), h23 = __floats2bfloat162_rn(f.z, f.w);
                const float2 b01 = __bfloat1622float2(h01), b23 = __bfloat1622float2(h23);
                const __nv_bfloat162 l01 = __floats2bfloat162_rn(f.x - b01.x, f.y - b01.y), l23 = __floats2bfloat162_rn(f.z - b23.x, f.w - b23.y);
                const uint32_t off = core_offset(r, set * 32 + 4 * c4, XA_LBO, XA_SBO);
                *reinterpret_cast<uint2*>(s.xa[0] + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                *reinterpret_cast<uint2*>(s.xa[1] + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
        };
        if constexpr (AFFINE) {
#pragma unroll 1
            for (int p = 0; p < TILE_M / 32; p += 2) {
                const int r0 = p * 32 + (threadIdx.x >> 3), r1 = r0 + 32;
                TapSet t0, t1;
                float4 v0[12], v1[12];
                read_taps(r0, t0);
                read_taps(r1, t1);
                gather_load_streaming(a.set_norm, t0, c4, v0);
                gather_load_streaming(a.set_norm, t1, c4, v1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = h ? r1 : r0;
                    float4 f[3];
                    float w_in[3];
                    gather_reduce_planes(h ? v1 : v0, h ? t1 : t0, f, w_in);
                    constexpr float third = 1.0f / 3.0f;
                    const float4 fa = make_float4(((f[0].x + f[1].x) + f[2].x) * third, ((f[0].y + f[1].y) + f[2].y) * third,
                                                  ((f[0].z + f[1].z) + f[2].z) * third, ((f[0].w + f[1].w) + f[2].w) * third);
                    const int64_t item = a.affine_items == 1 ? 0 : s.tap_item[r];
                    const float4* sc = reinterpret_cast<const float4*>(a.affine_scale + item * 96) + c4;
                    const float4* sh = reinterpret_cast<const float4*>(a.affine_shift + item * 96) + c4;
                    float4 dn[3];
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        const float4 scl = __ldg(sc + pl * 8), shf = __ldg(sh + pl * 8);
                        dn[pl] = make_float4(fmaf(scl.x, f[pl].x, shf.x * w_in[pl]), fmaf(scl.y, f[pl].y, shf.y * w_in[pl]),
                                             fmaf(scl.z, f[pl].z, shf.z * w_in[pl]), fmaf(scl.w, f[pl].w, shf.w * w_in[pl]));
                        __half2* dst = reinterpret_cast<__half2*>(&s.fp[r][pl * 32 + 4 * c4]);
                        dst[0] = __floats2half2_rn(f[pl].x, f[pl].y);
                        dst[1] = __floats2half2_rn(f[pl].z, f[pl].w);
                    }
                    const float4 fb = make_float4(((dn[0].x + dn[1].x) + dn[2].x) * third, ((dn[0].y + dn[1].y) + dn[2].y) * third,
                                                  ((dn[0].z + dn[1].z) + dn[2].z) * third, ((dn[0].w + dn[1].w) + dn[2].w) * third);
                    store_x(r, fa, fb);
                }
            }
        } else {
#pragma unroll 1
            for (int p = 0; p < TILE_M / 32; ++p) {
                const int r = p * 32 + (threadIdx.x >> 3);
                TapSet ts;
                float4 va[12], vb[12];
                read_taps(r, ts);
                if constexpr (DIS) {
                    gather_load(a.set_norm, ts, c4, va);
                    gather_load(a.set_raw, ts, c4, vb);
                    store_x(r, gather_reduce(va, ts), gather_reduce(vb, ts));
                } else {                                       // both nets read the raw-plane features
                    gather_load(a.set_raw, ts, c4, vb);
                    const float4 f = gather_reduce(vb, ts);
                    store_x(r, f, f);
                }
            }
        }
        tc::fence_async_smem();
        __syncthreads();
        BWD_MARK(0);

        // ---- G1: pre = X W1^T for both nets
        if (threadIdx.x < 32) {                    // warp 0, warp-uniform; the elected lane issues (tc::elect_one)
            const bool leader = tc::elect_one();
            tc::fence_after_sync();
            constexpr uint32_t id1 = tc::make_idesc_bf16(TILE_M, HIDDEN);
#pragma unroll
            for (int n = 0; n < 2; ++n)
                issue_gemm<true>(leader, tmem + C_PRE + 64 * n, s.xa[0] + 4 * n * XA_LBO, s.xa[1] + 4 * n * XA_LBO, XA_LBO, XA_SBO, s.w1[n][0], s.w1[n][1],
                                 B1_LBO, B1_SBO, FEAT, id1);
            if (leader) tc::mma_commit(&s.bar);
        }
        if (threadIdx.x >= TILE_M && tile + gridDim.x < n_tiles) tap_compute((tile + gridDim.x) * TILE_M, next_ts, next_item);
        // record gradients (and the saved colours) of this thread's row: requested now, consumed after the softplus loop
        const bool live = base + row < a.total;
        float4 pg[9], py[8];
        {
            const float4* g4 = reinterpret_cast<const float4*>(a.g_rec + (base + row) * 48);
            const float4* r4 = reinterpret_cast<const float4*>(a.rec + (base + row) * 48);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            // records are {sigma, seg[15], rgb[32]}: the colour net needs d rgb and the saved rgb, the other one d sigma / d seg
            const bool colour_net = DIS ? net == 1 : net == 0;
            if (!colour_net) {
#pragma unroll
                for (int i = 0; i < 4; ++i) pg[i] = live ? __ldg(g4 + i) : z4;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) { pg[i] = live ? __ldg(g4 + 4 + i) : z4; py[i] = live ? __ldg(r4 + 4 + i) : z4; }
                if (!DIS) pg[8] = live ? __ldg(g4) : z4;             // d sigma rides with the colours in `net` (33 outputs)
            }
        }
        tc::mbar_wait(&s.bar, phase); phase ^= 1;
        tc::fence_after_sync();
        BWD_MARK(1);

        // ---- epilogue 1: hidden activations and output-layer gradients of this thread's (row, net)
        {
#pragma unroll 1
            for (int q = 0; q < HIDDEN / 16; ++q) {
                float v[16];
                tc::tmem_ld16(lane_addr + C_PRE + 64 * net + q * 16, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = softplus_fast(v[i] + s.bias1[net][q * 16 + i]);
                float h8[8];
#pragma unroll
                for (int c8 = 0; c8 < 2; ++c8) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) h8[i] = v[c8 * 8 + i];
                    store8(s.hc[0], s.hc[1], core_offset(row, 64 * net + q * 16 + c8 * 8, HC_LBO, HC_SBO), h8);
                }
            }
            // colour gradients through rgb = s*1.002 - 0.001, s = sigmoid(out):  d rgb / d out = 1.002 * s * (1 - s), s from the saved rgb
            auto colour_dy = [&](int c) {
                const float4 g = pg[c >> 2], y = py[c >> 2];
                const int i = c & 3;
                const float gv = i == 0 ? g.x : i == 1 ? g.y : i == 2 ? g.z : g.w, yv = i == 0 ? y.x : i == 1 ? y.y : i == 2 ? y.z : y.w;
                const float sg = (yv + 0.001f) * (1.0f / 1.002f);
                return gv * 1.002f * sg * (1.0f - sg);
            };
            auto rec_g = [&](int c) {                                   // element c of the first 16 record gradients {d sigma, d seg[15]}
                const float4 g = pg[c >> 2];
                const int i = c & 3;
                return i == 0 ? g.x : i == 1 ? g.y : i == 2 ? g.z : g.w;
            };
            if constexpr (DIS) {
                if (net == 0) {                                         // geo_net: [d sigma, d seg[15]]
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        float d8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { d8[i] = rec_g(c8 * 8 + i); db2[c8 * 8 + i] += d8[i]; }
                        store8(s.dy[0], s.dy[1], core_offset(row, c8 * 8, DY_LBO, DY_SBO), d8);
                    }
                } else {                                                // app_net: 32 colours
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float d8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { d8[i] = colour_dy(c8 * 8 + i); db2[c8 * 8 + i] += d8[i]; }
                        store8(s.dy[0], s.dy[1], core_offset(row, O::PAD0 + c8 * 8, DY_LBO, DY_SBO), d8);
                    }
                }
            } else {
                if (net == 0) {                                         // net: [d sigma, 32 colours, 15 x 0]
#pragma unroll
                    for (int c8 = 0; c8 < O::PAD0 / 8; ++c8) {
                        float d8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int k = c8 * 8 + i;
                            d8[i] = k == 0 ? pg[8].x : (k <= 32 ? colour_dy(k - 1) : 0.0f);
                            db2[k] += d8[i];
                        }
                        store8(s.dy[0], s.dy[1], core_offset(row, c8 * 8, DY_LBO, DY_SBO), d8);
                    }
                } else {                                                // seg_net: [d seg[15], 0]; absent in OSG
#pragma unroll
                    for (int c8 = 0; c8 < O::PAD1 / 8; ++c8) {
                        float d8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int k = c8 * 8 + i;
                            d8[i] = (O::OUT1 > 0 && k < 15) ? rec_g(k + 1) : 0.0f;
                            db2[k] += d8[i];
                        }
                        store8(s.dy[0], s.dy[1], core_offset(row, O::PAD0 + c8 * 8, DY_LBO, DY_SBO), d8);
                    }
                }
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        BWD_MARK(2);

        // ---- G2: dH = dY W2;  G4: dW2^T += H^T dY (contraction over the tile rows)
        if (threadIdx.x < 32) {
            const bool leader = tc::elect_one();
            tc::fence_after_sync();
            constexpr uint32_t id2 = tc::make_idesc_bf16(TILE_M, HIDDEN);
            issue_gemm<true>(leader, tmem + C_DH, s.dy[0], s.dy[1], DY_LBO, DY_SBO, s.w2t[0][0], s.w2t[0][1], W2T_LBO, W2T_SBO, O::PAD0, id2);
            issue_gemm<true>(leader, tmem + C_DH + 64, s.dy[0] + (O::PAD0 / 8) * DY_LBO, s.dy[1] + (O::PAD0 / 8) * DY_LBO, DY_LBO, DY_SBO, s.w2t[1][0], s.w2t[1][1],
                             W2T_LBO, W2T_SBO, O::PAD1, id2);
            issue_gemm_rows(leader, tmem + C_DW2, s.hc[0], s.hc[1], HC_LBO, HC_SBO, s.dy[0], s.dy[1], DY_LBO, DY_SBO, idesc_mn(TILE_M, DY_COLS), !first);
            if (leader) tc::mma_commit(&s.bar);
        }
        tc::mbar_wait(&s.bar, phase); phase ^= 1;
        tc::fence_after_sync();
        BWD_MARK(3);

        // ---- epilogue 2: dpre = dH * sigmoid(pre + b1), over the hidden tile
#pragma unroll 1
        for (int q = 0; q < HIDDEN / 16; ++q) {
            float pre[16], dh[16];
            tc::tmem_ld16(lane_addr + C_PRE + 64 * net + q * 16, pre);
            tc::tmem_ld16(lane_addr + C_DH + 64 * net + q * 16, dh);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) dh[i] *= sigmoid_fast(pre[i] + s.bias1[net][q * 16 + i]);
            float d8[8];
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) d8[i] = dh[c8 * 8 + i];
                store8(s.hc[0], s.hc[1], core_offset(row, 64 * net + q * 16 + c8 * 8, HC_LBO, HC_SBO), d8);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        BWD_MARK(4);

        // ---- G3: dX = dpre W1;  G5: [dW1 | db1] += dpre^T [X | 1]
        if (threadIdx.x < 32) {
            const bool leader = tc::elect_one();
            tc::fence_after_sync();
            constexpr uint32_t id3 = tc::make_idesc_bf16(TILE_M, FEAT);
#pragma unroll
            for (int n = 0; n < 2; ++n)
                issue_gemm<true>(leader, tmem + C_DX + 32 * n, s.hc[0] + 8 * n * HC_LBO, s.hc[1] + 8 * n * HC_LBO, HC_LBO, HC_SBO, s.w1t[n][0], s.w1t[n][1],
                                 W1T_LBO, W1T_SBO, HIDDEN, id3);
            issue_gemm_rows(leader, tmem + C_DW1, s.hc[0], s.hc[1], HC_LBO, HC_SBO, s.xa[0], s.xa[1], XA_LBO, XA_SBO, idesc_mn(TILE_M, XA_COLS), !first);
            if (leader) tc::mma_commit(&s.bar);
        }
        tc::mbar_wait(&s.bar, phase); phase ^= 1;
        tc::fence_after_sync();
        first = false;
        BWD_MARK(5);

        // ---- epilogue 3: dX of this (row, set) -> fp32 staging (over the hidden tile, which G5 has finished reading)
        float* gx = reinterpret_cast<float*>(s.hc[0]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float v[16];
            tc::tmem_ld16(lane_addr + C_DX + 32 * net + q * 16, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(gx + row * GX_STRIDE + 32 * net + 16 * q + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        tc::fence_before_sync();
        __syncthreads();
        BWD_MARK(6);
        // ---- gather backward: w_tap/3 * dX into the channel-last plane gradients
        constexpr float third = 1.0f / 3.0f;
#pragma unroll 1
        for (int p = 0; p < TILE_M / 32; ++p) {
            const int r = p * 32 + (threadIdx.x >> 3);
            if (base + r >= a.total) continue;
            float4 gn = *reinterpret_cast<const float4*>(gx + r * GX_STRIDE + 4 * c4), gr = *reinterpret_cast<const float4*>(gx + r * GX_STRIDE + 32 + 4 * c4);
            gn = make_float4(gn.x * third, gn.y * third, gn.z * third, gn.w * third);
            gr = make_float4(gr.x * third, gr.y * third, gr.z * third, gr.w * third);
            if constexpr (AFFINE) {
                const int64_t item = a.affine_items == 1 ? 0 : s.tap_item[r];
                const float4* sc = reinterpret_cast<const float4*>(a.affine_scale + item * 96) + c4;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const float4 scl = __ldg(sc + p * 8);
                    const float4 gf = make_float4(fmaf(scl.x, gr.x, gn.x), fmaf(scl.y, gr.y, gn.y), fmaf(scl.z, gr.z, gn.z), fmaf(scl.w, gr.w, gn.w));
                    float w_in = 0.0f;
                    const float4 w4 = *reinterpret_cast<const float4*>(&s.tap_w[r][4 * p]);      // one LDS.128 each for the plane's 4 taps
                    const int4 o4 = *reinterpret_cast<const int4*>(&s.tap_off[r][4 * p]);
                    const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
                    const int os[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float w = ws[k];
                        w_in += w;
                        if (w != 0.0f)
                            red_add_v4(a.g_norm + (int64_t)os[k] * 4 + 4 * c4, make_float4(gf.x * w, gf.y * w, gf.z * w, gf.w * w));
                    }
                    const __half2* src = reinterpret_cast<const __half2*>(&s.fp[r][p * 32 + 4 * c4]);
                    const float2 f01 = __half22float2(src[0]), f23 = __half22float2(src[1]);
                    const float4 dsv = make_float4(f01.x * gr.x, f01.y * gr.y, f23.x * gr.z, f23.y * gr.w);
                    const float4 dmv = make_float4(w_in * gr.x, w_in * gr.y, w_in * gr.z, w_in * gr.w);
                    if (!straddle) {
                        ds[p].x += dsv.x; ds[p].y += dsv.y; ds[p].z += dsv.z; ds[p].w += dsv.w;
                        dm[p].x += dmv.x; dm[p].y += dmv.y; dm[p].z += dmv.z; dm[p].w += dmv.w;
                    } else {
                        float* gs = a.g_scale + item * 96 + p * 32 + 4 * c4;
                        float* gm = a.g_shift + item * 96 + p * 32 + 4 * c4;
                        atomicAdd(gs, dsv.x); atomicAdd(gs + 1, dsv.y); atomicAdd(gs + 2, dsv.z); atomicAdd(gs + 3, dsv.w);
                        atomicAdd(gm, dmv.x); atomicAdd(gm + 1, dmv.y); atomicAdd(gm + 2, dmv.z); atomicAdd(gm + 3, dmv.w);
                    }
                }
                continue;
            }
            if constexpr (!DIS) gr = make_float4(gn.x + gr.x, gn.y + gr.y, gn.z + gr.z, gn.w + gr.w);      // both nets read the raw planes
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float w = s.tap_w[r][i];
                if (w != 0.0f) {
                    const int64_t off = (int64_t)s.tap_off[r][i] * 4 + 4 * c4;
                    if constexpr (DIS) red_add_v4(a.g_norm + off, make_float4(gn.x * w, gn.y * w, gn.z * w, gn.w * w));
                    red_add_v4(a.g_raw + off, make_float4(gr.x * w, gr.y * w, gr.z * w, gr.w * w));
                }
            }
        }
        __syncthreads();      // the staging tile and the taps are reused by the next tile
        BWD_MARK(7);
    }
#ifdef NFE_BWD_PROFILE
    if (threadIdx.x == 0)
        for (int i = 0; i < 16; ++i) atomicAdd(&g_bwd_prof[i], (unsigned long long)prof_[i]);
#endif

    flush_stats();
    // ---- parameter gradients of this CTA -> global (chain rule through the FullyConnectedLayer gains)
    if (!first) {
        tc::fence_after_sync();
        if (warp < 4 && ((row >> 6) == 0 || O::OUT1 > 0)) {     // (OSG: the rows of the absent second net carry nothing)
            const int n = row >> 6, j = row & 63;               // accumulator row = hidden unit j of net n
            const nfe_mlp& p = n ? app : geo;
            const uint32_t la = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int q = 0; q < XA_COLS / 16; ++q) {
                float v[16];
                tc::tmem_ld16(la + C_DW1 + q * 16, v);
                tc::tmem_ld_wait();
                if (q == 4) atomicAdd(a.gb1[n] + j, v[0] * p.bgain1);
                else if ((q >> 1) == n) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(a.gw1[n] + j * FEAT + (q & 1) * 16 + i, v[i] * p.wgain1);
                }
            }
#pragma unroll 1
            for (int q = 0; q < DY_COLS / 16; ++q) {
                float v[16];
                tc::tmem_ld16(la + C_DW2 + q * 16, v);
                tc::tmem_ld_wait();
                // columns [0, PAD0) belong to net 0, [PAD0, PAD0 + PAD1) to net 1; only the real outputs of this row's net are flushed
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = q * 16 + i;
                    const int o = n == 0 ? c : c - O::PAD0;
                    if ((n == 0 ? c < O::PAD0 : c >= O::PAD0) && o < (n == 0 ? O::OUT0 : O::OUT1)) atomicAdd(a.gw2[n] + o * HIDDEN + j, v[i] * p.wgain2);
                }
            }
        }
        // db2: column sums over the rows this thread handled, reduced over the warp
        const int n_out = net ? O::OUT1 : O::OUT0;
#pragma unroll
        for (int i = 0; i < DB2; ++i) {
            float v = db2[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && i < n_out) atomicAdd(a.gb2[net] + i, v * mine.bgain2);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TMEM_ALLOC);
}

}  // namespace fb
}  // namespace nfe

using namespace nfe;

// shared by nfe_field_bwd (samples along rays) and nfe_run_model_bwd (explicit points: coords != NULL, s_per_ray = 1)
static int field_bwd_launch(const char* who, int kind, const float* planes_norm_cl, const float* planes_cl, int plane_batch, int height, int width,
                            float box_warp, const float* origins, const float* dirs, const float* depths, const float* coords, int n,
                            int64_t n_rays, int s_per_ray, const nfe_mlp* net_a, const nfe_mlp* net_b, const float* rec, const float* g_rec,
                            float* g_planes_norm_cl, float* g_planes_cl, float* g_w1_a, float* g_b1_a, float* g_w2_a, float* g_b2_a, float* g_w1_b,
                            float* g_b1_b, float* g_w2_b, float* g_b2_b, const float* affine_scale, const float* affine_shift, int affine_items,
                            float* g_affine_scale, float* g_affine_shift, nfe_stream_t stream)
{
    const int64_t total = (int64_t)n * n_rays * s_per_ray;
    if (total == 0) return 0;
    const bool affine = affine_scale != nullptr;
    const bool dis = kind == NFE_DEC_DISENTANGLED, osg = kind == NFE_DEC_OSG;
    NFE_REQUIRE(kind == NFE_DEC_DISENTANGLED || kind == NFE_DEC_SEGMENTATION || kind == NFE_DEC_OSG, "%s: unknown decoder kind %d", who, kind);
    NFE_REQUIRE((coords || (origins && dirs && depths)) && net_a && (net_b || osg) && rec && g_rec, "%s: null pointer", who);
    NFE_REQUIRE(!dis || (planes_norm_cl && g_planes_norm_cl), "%s: the disentangled decoder needs the normalised planes and their gradient buffer", who);
    if (affine) {
        NFE_REQUIRE(dis, "%s: the single-gather backward belongs to the disentangled decoder", who);
        NFE_REQUIRE(affine_shift && g_affine_scale && g_affine_shift, "%s: the single-gather backward needs shift and both statistics gradients", who);
        NFE_REQUIRE(affine_items == n || affine_items == 1, "%s: %d statistics rows for a batch of %d", who, affine_items, n);
    } else {
        NFE_REQUIRE(planes_cl && g_planes_cl, "%s: null raw-plane pointer (and no affine statistics)", who);
    }
    NFE_REQUIRE(g_w1_a && g_b1_a && g_w2_a && g_b2_a && (osg || (g_w1_b && g_b1_b && g_w2_b && g_b2_b)), "%s: null parameter-gradient pointer", who);
    const int out_a = dis ? 16 : 33, out_b = dis ? 32 : 15;
    NFE_REQUIRE(net_a->in_dim == FEAT && net_a->hidden == HIDDEN && net_a->out_dim == out_a &&
                (osg || (net_b->in_dim == FEAT && net_b->hidden == HIDDEN && net_b->out_dim == out_b)),
                "%s: decoder widths must be 32-64-%d%s", who, out_a, osg ? "" : (dis ? " / 32-64-32" : " / 32-64-15"));
    NFE_REQUIRE(plane_batch == n || plane_batch == 1, "%s: plane batch %d does not match batch %d", who, plane_batch, n);
    NFE_REQUIRE((int64_t)plane_batch * height * width * 3 * (FEAT / 4) < (1ll << 31), "%s: planes exceed the 32-bit texel offsets", who);
    NFE_REQUIRE(box_warp != 0.0f, "%s: box_warp must be non-zero", who);
    fb::Args a = {};
    a.set_norm = planes_norm_cl; a.set_raw = planes_cl; a.g_norm = g_planes_norm_cl; a.g_raw = g_planes_cl;
    a.plane_batch = plane_batch; a.H = height; a.W = width; a.scale = (float)(2.0 / (double)box_warp);
    a.origins = origins; a.dirs = dirs; a.depths = depths; a.coords = coords; a.s_per_ray = s_per_ray; a.m = n_rays * s_per_ray; a.total = total;
    a.rec = rec; a.g_rec = g_rec;
    a.gw1[0] = g_w1_a; a.gb1[0] = g_b1_a; a.gw2[0] = g_w2_a; a.gb2[0] = g_b2_a;
    a.gw1[1] = g_w1_b; a.gb1[1] = g_b1_b; a.gw2[1] = g_w2_b; a.gb2[1] = g_b2_b;
    a.affine_scale = affine_scale; a.affine_shift = affine_shift; a.affine_items = affine_items;
    a.g_scale = g_affine_scale; a.g_shift = g_affine_shift;
    const nfe_mlp none = {};
    const nfe_mlp& nb = net_b ? *net_b : none;
    const int64_t n_tiles = (total + TILE_M - 1) / TILE_M;
    const int64_t cap = sm_count();
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    auto launch = [&](auto kernel, size_t smem, bool& configured) -> int {
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            NFE_REQUIRE(e == cudaSuccess, "%s: cannot reserve %zu bytes of shared memory: %s", who, smem, cudaGetErrorString(e));
            configured = true;
        }
        kernel<<<grid, fb::THREADS, smem, as_stream(stream)>>>(a, *net_a, nb);
        return 0;
    };
    static bool configured[4] = {false, false, false, false};
    int rc;
    if (dis && affine) rc = launch(fb::field_bwd_kernel<NFE_DEC_DISENTANGLED, true>, sizeof(fb::Smem<NFE_DEC_DISENTANGLED, true>) + 128, configured[0]);
    else if (dis) rc = launch(fb::field_bwd_kernel<NFE_DEC_DISENTANGLED, false>, sizeof(fb::Smem<NFE_DEC_DISENTANGLED, false>) + 128, configured[1]);
    else if (osg) rc = launch(fb::field_bwd_kernel<NFE_DEC_OSG, false>, sizeof(fb::Smem<NFE_DEC_OSG, false>) + 128, configured[2]);
    else rc = launch(fb::field_bwd_kernel<NFE_DEC_SEGMENTATION, false>, sizeof(fb::Smem<NFE_DEC_SEGMENTATION, false>) + 128, configured[3]);
    if (rc) return rc;
    NFE_LAUNCH_CHECK("field_bwd_kernel");
    return 0;
}

NFE_EXPORT int nfe_field_bwd(int kind, const float* planes_norm_cl, const float* planes_cl, int plane_batch, int height, int width, float box_warp,
                             const float* origins, const float* dirs, const float* depths, int n, int64_t n_rays, int s_per_ray,
                             const nfe_mlp* net_a, const nfe_mlp* net_b, const float* rec, const float* g_rec, float* g_planes_norm_cl,
                             float* g_planes_cl, float* g_w1_a, float* g_b1_a, float* g_w2_a, float* g_b2_a, float* g_w1_b, float* g_b1_b,
                             float* g_w2_b, float* g_b2_b, const float* affine_scale, const float* affine_shift, int affine_items,
                             float* g_affine_scale, float* g_affine_shift, nfe_stream_t stream)
{
    return field_bwd_launch("nfe_field_bwd", kind, planes_norm_cl, planes_cl, plane_batch, height, width, box_warp, origins, dirs, depths, nullptr, n,
                            n_rays, s_per_ray, net_a, net_b, rec, g_rec, g_planes_norm_cl, g_planes_cl, g_w1_a, g_b1_a, g_w2_a, g_b2_a, g_w1_b, g_b1_b,
                            g_w2_b, g_b2_b, affine_scale, affine_shift, affine_items, g_affine_scale, g_affine_shift, stream);
}

NFE_EXPORT int nfe_run_model_bwd(int kind, const float* planes_norm_cl, const float* planes_cl, int plane_batch, int height, int width, float box_warp,
                                 const float* coords, int n, int64_t m, const nfe_mlp* net_a, const nfe_mlp* net_b, const float* rec,
                                 const float* g_rec, float* g_planes_norm_cl, float* g_planes_cl, float* g_w1_a, float* g_b1_a, float* g_w2_a,
                                 float* g_b2_a, float* g_w1_b, float* g_b1_b, float* g_w2_b, float* g_b2_b, const float* affine_scale,
                                 const float* affine_shift, int affine_items, float* g_affine_scale, float* g_affine_shift, nfe_stream_t stream)
{
    NFE_REQUIRE(coords || (int64_t)n * m == 0, "nfe_run_model_bwd: null coords");
    return field_bwd_launch("nfe_run_model_bwd", kind, planes_norm_cl, planes_cl, plane_batch, height, width, box_warp, nullptr, nullptr, nullptr, coords,
                            n, m, 1, net_a, net_b, rec, g_rec, g_planes_norm_cl, g_planes_cl, g_w1_a, g_b1_a, g_w2_a, g_b2_a, g_w1_b, g_b1_b, g_w2_b,
                            g_b2_b, affine_scale, affine_shift, affine_items, g_affine_scale, g_affine_shift, stream);
}

#ifdef NFE_BWD_PROFILE
NFE_EXPORT int nfe_debug_bwd_profile(unsigned long long* out16, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, nfe::fb::g_bwd_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(nfe::fb::g_bwd_prof, z, sizeof(z)); }
    return 0;
}
#endif
