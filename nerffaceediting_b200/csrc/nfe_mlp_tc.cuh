// Decoder MLPs on the 5th-generation tensor cores: one 128-sample tile per CTA step.
//
//   layer 1   D1[net] (128x64, TMEM fp32) = X[set(net)] (128x32 bf16, smem) * W1[net]^T (64x32 bf16, smem)
//   epilogue  H = softplus(D1 + b1) -> bf16 A operand in smem (thread m owns row m = TMEM lane m)
//   layer 2   D2[net] (128xN, TMEM fp32) = H[net] (128x64) * W2[net]^T (Nx64)
//   epilogue  + b2, sigmoid clamp for colour, results to the caller
//
// Precision: SPLIT = true runs every product as three bf16 MMAs (hi*hi + lo*hi + hi*lo with
// x = hi + lo, fp32 accumulation in TMEM), which keeps ~16 significand bits per operand and meets
// the fp32 tolerance of the path (1e-4); SPLIT = false is plain bf16 (1e-2 tolerance).
#pragma once
#include "nfe_field.cuh"
#include "nfe_tc.cuh"

namespace nfe { namespace tcmlp {

constexpr int TILE_M = 128;
// feature tile (128 x 32): core matrices padded to 160 B along K so that the gather's 8-byte
// stores (8 lanes per sample row) are bank-conflict free
constexpr int A1_LBO = 160, A1_SBO = 640, A1_BYTES = 16 * A1_SBO;
constexpr int B1_LBO = 128, B1_SBO = 512, B1_BYTES = 8 * B1_SBO;       // W1: 64 x 32
constexpr int A2_LBO = 128, A2_SBO = 1024, A2_BYTES = 16 * A2_SBO;     // hidden: 128 x 64
constexpr int B2_LBO = 128, B2_SBO = 1024;                             // W2: N x 64 -> (N/8) * 1024 bytes

template <int KIND> struct TcTraits;
template <> struct TcTraits<NFE_DEC_OSG> { static constexpr int N_A = 48, N_B = 16, OUT_A = 33, OUT_B = 0, SETS = 1; static constexpr bool HAS_B = false; };
template <> struct TcTraits<NFE_DEC_DISENTANGLED> { static constexpr int N_A = 16, N_B = 32, OUT_A = 16, OUT_B = 32, SETS = 2; static constexpr bool HAS_B = true; };
template <> struct TcTraits<NFE_DEC_SEGMENTATION> { static constexpr int N_A = 48, N_B = 16, OUT_A = 33, OUT_B = 15, SETS = 1; static constexpr bool HAS_B = true; };

// TMEM columns (fp32): D1 of net A / B, then D2 of net A / B
constexpr int COL_D1A = 0, COL_D1B = 64, COL_D2A = 128, COL_D2B = 176, TMEM_COLS = 256;

template <int KIND, bool SPLIT>
struct Smem {
    using T = TcTraits<KIND>;
    static constexpr int PARTS = SPLIT ? 2 : 1;
    static constexpr int NETS = T::HAS_B ? 2 : 1;
    alignas(128) unsigned char a1[T::SETS][PARTS][A1_BYTES];
    alignas(128) unsigned char b1[NETS][PARTS][B1_BYTES];
    alignas(128) unsigned char a2[NETS][PARTS][A2_BYTES];
    alignas(128) unsigned char b2a[PARTS][(T::N_A / 8) * B2_SBO];
    alignas(128) unsigned char b2b[PARTS][(T::N_B / 8) * B2_SBO];
    float bias1[NETS][HIDDEN];
    float bias2a[T::N_A];
    float bias2b[T::N_B];
    alignas(8) uint64_t bar[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t core_offset(int row, int k, int lbo, int sbo)
{
    return (uint32_t)((row >> 3) * sbo + (k >> 3) * lbo + (row & 7) * 16 + (k & 7) * 2);
}

// Block-cooperative: gain-folded weights (networks_stylegan2.py:115-120) split into bf16 parts and
// laid out as K-major B operands; rows beyond out_dim are zero.
template <int PARTS>
__device__ void load_weights(unsigned char* dst, size_t part_stride, const float* w, float gain, int rows, int rows_pad, int K, int lbo, int sbo)
{
    for (int i = threadIdx.x; i < rows_pad * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        const float v = n < rows ? __fmul_rn(__ldg(w + n * K + k), gain) : 0.0f;
        __nv_bfloat16 hi, lo;
        tc::split_bf16(v, hi, lo);
        const uint32_t off = core_offset(n, k, lbo, sbo);
        *reinterpret_cast<__nv_bfloat16*>(dst + off) = hi;
        if (PARTS == 2) *reinterpret_cast<__nv_bfloat16*>(dst + part_stride + off) = lo;
    }
}

__device__ __forceinline__ float folded_bias(const float* b, float gain, int i) { return gain != 1.0f ? __fmul_rn(__ldg(b + i), gain) : __ldg(b + i); }

template <int KIND, bool SPLIT>
__device__ void load_params(Smem<KIND, SPLIT>& s, const nfe_mlp& net_a, const nfe_mlp& net_b)
{
    using T = TcTraits<KIND>;
    constexpr int PARTS = SPLIT ? 2 : 1;
    load_weights<PARTS>(s.b1[0][0], B1_BYTES, net_a.w1, net_a.wgain1, HIDDEN, HIDDEN, FEAT, B1_LBO, B1_SBO);
    load_weights<PARTS>(s.b2a[0], sizeof(s.b2a[0]), net_a.w2, net_a.wgain2, T::OUT_A, T::N_A, HIDDEN, B2_LBO, B2_SBO);
    for (int i = threadIdx.x; i < HIDDEN; i += blockDim.x) s.bias1[0][i] = folded_bias(net_a.b1, net_a.bgain1, i);
    for (int i = threadIdx.x; i < T::N_A; i += blockDim.x) s.bias2a[i] = i < T::OUT_A ? folded_bias(net_a.b2, net_a.bgain2, i) : 0.0f;
    if constexpr (T::HAS_B) {
        load_weights<PARTS>(s.b1[1][0], B1_BYTES, net_b.w1, net_b.wgain1, HIDDEN, HIDDEN, FEAT, B1_LBO, B1_SBO);
        load_weights<PARTS>(s.b2b[0], sizeof(s.b2b[0]), net_b.w2, net_b.wgain2, T::OUT_B, T::N_B, HIDDEN, B2_LBO, B2_SBO);
        for (int i = threadIdx.x; i < HIDDEN; i += blockDim.x) s.bias1[1][i] = folded_bias(net_b.b1, net_b.bgain1, i);
        for (int i = threadIdx.x; i < T::N_B; i += blockDim.x) s.bias2b[i] = i < T::OUT_B ? folded_bias(net_b.b2, net_b.bgain2, i) : 0.0f;
    }
}

// Row `row` of a feature set: 4 channels starting at k0, as bf16 parts (8-byte stores).
template <bool SPLIT>
__device__ __forceinline__ void store_features4(unsigned char (*a1)[A1_BYTES], int row, int k0, float4 f)
{
    // packed conversions (F2FP, two values per instruction); the scalar form goes through the XU pipe the softplus needs
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(f.x, f.y), h23 = __floats2bfloat162_rn(f.z, f.w);
    const uint32_t off = core_offset(row, k0, A1_LBO, A1_SBO);
    *reinterpret_cast<uint2*>(a1[0] + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    if (SPLIT) {
        const float2 neg2 = make_float2(-1.0f, -1.0f);
        const float2 d01 = ffma2(__bfloat1622float2(h01), neg2, make_float2(f.x, f.y)), d23 = ffma2(__bfloat1622float2(h23), neg2, make_float2(f.z, f.w));
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(d01.x, d01.y), l23 = __floats2bfloat162_rn(d23.x, d23.y);
        *reinterpret_cast<uint2*>(a1[1] + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
}

// A whole GEMM: D (+)= A * B^T over K (multiple of 16), with the three split terms.  Called by every lane of the issuing warp
// (warp-uniform arguments); only `leader` (tc::elect_one()) issues.  The descriptors differ in their address field alone, so each is the
// constant part plus an add.  The one-argument-less form is for callers that are a single thread already.
template <bool SPLIT>
__device__ __forceinline__ void issue_gemm(bool leader, uint32_t tmem_d, const unsigned char* a_hi, const unsigned char* a_lo, int a_lbo, int a_sbo,
                                           const unsigned char* b_hi, const unsigned char* b_lo, int b_lbo, int b_sbo, int K, uint32_t idesc)
{
    const uint64_t da0 = tc::make_desc(0, a_lbo, a_sbo), db0 = tc::make_desc(0, b_lbo, b_sbo);
    const uint32_t ah = tc::smem_u32(a_hi) >> 4, al = tc::smem_u32(a_lo) >> 4, bh = tc::smem_u32(b_hi) >> 4, bl = tc::smem_u32(b_lo) >> 4;
    const uint32_t ak = (uint32_t)(2 * a_lbo) >> 4, bk = (uint32_t)(2 * b_lbo) >> 4;
    uint32_t acc = 0;
    constexpr int TERMS = SPLIT ? 3 : 1;
#pragma unroll
    for (int t = 0; t < TERMS; ++t) {
        const uint32_t a = (t == 1) ? al : ah, b = (t == 2) ? bl : bh;   // hi*hi, lo*hi, hi*lo
#pragma unroll 4
        for (int k = 0; k < (K >> 4); ++k) {
            if (leader) tc::mma_bf16_ss(tmem_d, da0 + (a + k * ak), db0 + (b + k * bk), idesc, acc);
            acc = 1;
        }
    }
}
template <bool SPLIT>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const unsigned char* a_hi, const unsigned char* a_lo, int a_lbo, int a_sbo,
                                           const unsigned char* b_hi, const unsigned char* b_lo, int b_lbo, int b_sbo, int K, uint32_t idesc)
{
    issue_gemm<SPLIT>(true, tmem_d, a_hi, a_lo, a_lbo, a_sbo, b_hi, b_lo, b_lbo, b_sbo, K, idesc);
}

template <int KIND, bool SPLIT>
__device__ __forceinline__ void issue_layer1(Smem<KIND, SPLIT>& s, uint32_t tmem)
{
    using T = TcTraits<KIND>;
    constexpr int P = SPLIT ? 1 : 0;
    constexpr uint32_t idesc = tc::make_idesc_bf16(TILE_M, HIDDEN);
    // net A reads the first feature set (normalised planes for the disentangled decoder, the only set otherwise)
    issue_gemm<SPLIT>(tmem + COL_D1A, s.a1[0][0], s.a1[0][P], A1_LBO, A1_SBO, s.b1[0][0], s.b1[0][P], B1_LBO, B1_SBO, FEAT, idesc);
    if constexpr (T::HAS_B) {
        constexpr int SET_B = T::SETS - 1;
        issue_gemm<SPLIT>(tmem + COL_D1B, s.a1[SET_B][0], s.a1[SET_B][P], A1_LBO, A1_SBO, s.b1[1][0], s.b1[1][P], B1_LBO, B1_SBO, FEAT, idesc);
    }
}

template <int KIND, bool SPLIT>
__device__ __forceinline__ void issue_layer2(Smem<KIND, SPLIT>& s, uint32_t tmem)
{
    using T = TcTraits<KIND>;
    constexpr int P = SPLIT ? 1 : 0;
    issue_gemm<SPLIT>(tmem + COL_D2A, s.a2[0][0], s.a2[0][P], A2_LBO, A2_SBO, s.b2a[0], s.b2a[P], B2_LBO, B2_SBO, HIDDEN,
                      tc::make_idesc_bf16(TILE_M, T::N_A));
    if constexpr (T::HAS_B)
        issue_gemm<SPLIT>(tmem + COL_D2B, s.a2[1][0], s.a2[1][P], A2_LBO, A2_SBO, s.b2b[0], s.b2b[P], B2_LBO, B2_SBO, HIDDEN,
                          tc::make_idesc_bf16(TILE_M, T::N_B));
}

// Epilogue 1 for row `row` (= TMEM lane): hidden = softplus(D1 + b1) -> bf16 parts of the layer-2 A operand.
template <bool SPLIT>
__device__ __forceinline__ void hidden_epilogue(uint32_t taddr_row, const float* bias1, unsigned char (*a2)[A2_BYTES], int row)
{
#pragma unroll
    for (int q = 0; q < HIDDEN / 16; ++q) {
        float v[16];
        tc::tmem_ld16(taddr_row + q * 16, v);
        tc::tmem_ld_wait();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float h0 = softplus_fast(v[2 * i] + bias1[q * 16 + 2 * i]);
            const float h1 = softplus_fast(v[2 * i + 1] + bias1[q * 16 + 2 * i + 1]);
            __nv_bfloat16 a, b, c, d;
            tc::split_bf16(h0, a, c); tc::split_bf16(h1, b, d);
            hi[i] = tc::pack_bf16(a, b); lo[i] = tc::pack_bf16(c, d);
        }
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) {
            const uint32_t off = core_offset(row, q * 16 + c8 * 8, A2_LBO, A2_SBO);
            *reinterpret_cast<uint4*>(a2[0] + off) = make_uint4(hi[4 * c8], hi[4 * c8 + 1], hi[4 * c8 + 2], hi[4 * c8 + 3]);
            if (SPLIT) *reinterpret_cast<uint4*>(a2[1] + off) = make_uint4(lo[4 * c8], lo[4 * c8 + 1], lo[4 * c8 + 2], lo[4 * c8 + 3]);
        }
    }
}

// Epilogue 2: this row's decoder outputs (sigma, 15 semantic logits, 32 colours) from D2 + b2.
template <int KIND, bool SPLIT>
__device__ __forceinline__ void output_epilogue(const Smem<KIND, SPLIT>& s, uint32_t taddr_row, float& sigma, float (&segv)[16], float (&col)[32])
{
    using T = TcTraits<KIND>;
    float a[T::N_A];
#pragma unroll
    for (int q = 0; q < T::N_A / 16; ++q) {
        float v[16];
        tc::tmem_ld16(taddr_row + COL_D2A + q * 16, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) a[q * 16 + i] = v[i] + s.bias2a[q * 16 + i];
    }
    sigma = a[0];
    if constexpr (KIND == NFE_DEC_DISENTANGLED) {
#pragma unroll
        for (int c = 0; c < 15; ++c) segv[c] = a[1 + c];
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) col[c] = rgb_activation(a[1 + c]);
    }
    if constexpr (T::HAS_B) {
        float b[T::N_B];
#pragma unroll
        for (int q = 0; q < T::N_B / 16; ++q) {
            float v[16];
            tc::tmem_ld16(taddr_row + COL_D2B + q * 16, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) b[q * 16 + i] = v[i] + s.bias2b[q * 16 + i];
        }
        if constexpr (KIND == NFE_DEC_DISENTANGLED) {
#pragma unroll
            for (int c = 0; c < 32; ++c) col[c] = rgb_activation(b[c]);
        } else {
#pragma unroll
            for (int c = 0; c < 15; ++c) segv[c] = b[c];
        }
    }
}

}}  // namespace nfe::tcmlp
