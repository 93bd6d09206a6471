// Device-side building blocks of the radiance-field evaluation: channel-last tri-plane gather and
// the decoder MLPs (fp32 SIMT path).  Shared by the stage kernels (nfe_field.cu) and the fused
// renderer (nfe_render.cu).
#pragma once
#include "nfe_common.cuh"

namespace nfe {

constexpr int FEAT = 32;    // n_features of every decoder (triplane.py:48-50: 32 channels per plane)
constexpr int HIDDEN = 64;  // self.hidden_dim (triplane.py:170,195,235)

// Effective (gain-folded) parameters of one MLP in shared memory.
//   w1  [HIDDEN][FEAT]      row-major, read as broadcast float4 over k
//   b1  [HIDDEN]
//   w2t [HIDDEN][OUT_PAD]   transposed so that one hidden unit updates all outputs with float4 reads
//   b2  [OUT_PAD]
template <int OUT_PAD>
struct MlpParams {
    float w1[HIDDEN * FEAT];
    float b1[HIDDEN];
    float w2t[HIDDEN * OUT_PAD];
    float b2[OUT_PAD];
};

constexpr int pad4(int x) { return (x + 3) / 4 * 4; }

// Block-cooperative: fold the FullyConnectedLayer gains (networks_stylegan2.py:115-120:
// w = weight*weight_gain, b = bias*bias_gain only when the gain is not 1) and lay the
// parameters out for the kernel.  Re-read from global on every launch: training updates them.
template <int OUT_PAD>
__device__ void load_mlp(MlpParams<OUT_PAD>& dst, const nfe_mlp& src)
{
    for (int i = threadIdx.x; i < HIDDEN * FEAT; i += blockDim.x) dst.w1[i] = __fmul_rn(__ldg(src.w1 + i), src.wgain1);
    for (int i = threadIdx.x; i < HIDDEN; i += blockDim.x)
        dst.b1[i] = src.bgain1 != 1.0f ? __fmul_rn(__ldg(src.b1 + i), src.bgain1) : __ldg(src.b1 + i);
    for (int i = threadIdx.x; i < HIDDEN * OUT_PAD; i += blockDim.x) {
        const int j = i / OUT_PAD, o = i % OUT_PAD;
        dst.w2t[i] = o < src.out_dim ? __fmul_rn(__ldg(src.w2 + o * HIDDEN + j), src.wgain2) : 0.0f;
    }
    for (int i = threadIdx.x; i < OUT_PAD; i += blockDim.x)
        dst.b2[i] = i < src.out_dim ? (src.bgain2 != 1.0f ? __fmul_rn(__ldg(src.b2 + i), src.bgain2) : __ldg(src.b2 + i)) : 0.0f;
}

// Softplus for the hidden layer: max(x,0) + log(1 + exp(-|x|)) on the SFU (ex2/lg2.approx).
// Absolute error ~1e-7 (threshold-20 branch of torch.nn.Softplus differs from this by < 3e-9).
__device__ __forceinline__ float softplus_fast(float x)
{
    const float e = __expf(-fabsf(x));
    return fmaxf(x, 0.0f) + __logf(1.0f + e);
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// One sample per lane: out = b2 + W2 * softplus(b1 + W1 * x).  The hidden loop is kept rolled
// (code size); k and o loops are unrolled so x[] / out[] stay in registers.
template <int OUT_PAD>
__device__ __forceinline__ void mlp_eval(const MlpParams<OUT_PAD>& p, const float (&x)[FEAT], float (&out)[OUT_PAD])
{
#pragma unroll
    for (int o = 0; o < OUT_PAD; ++o) out[o] = p.b2[o];
#pragma unroll 2
    for (int j = 0; j < HIDDEN; ++j) {
        const float4* w = reinterpret_cast<const float4*>(p.w1 + j * FEAT);
        float a0 = p.b1[j], a1 = 0.0f;
#pragma unroll
        for (int k = 0; k < FEAT / 4; k += 2) {
            const float4 wa = w[k], wb = w[k + 1];
            a0 = fmaf(x[4 * k + 0], wa.x, a0); a0 = fmaf(x[4 * k + 1], wa.y, a0);
            a0 = fmaf(x[4 * k + 2], wa.z, a0); a0 = fmaf(x[4 * k + 3], wa.w, a0);
            a1 = fmaf(x[4 * k + 4], wb.x, a1); a1 = fmaf(x[4 * k + 5], wb.y, a1);
            a1 = fmaf(x[4 * k + 6], wb.z, a1); a1 = fmaf(x[4 * k + 7], wb.w, a1);
        }
        const float h = softplus_fast(a0 + a1);
        const float4* w2 = reinterpret_cast<const float4*>(p.w2t + j * OUT_PAD);
#pragma unroll
        for (int o = 0; o < OUT_PAD / 4; ++o) {
            const float4 v = w2[o];
            out[4 * o + 0] = fmaf(h, v.x, out[4 * o + 0]); out[4 * o + 1] = fmaf(h, v.y, out[4 * o + 1]);
            out[4 * o + 2] = fmaf(h, v.z, out[4 * o + 2]); out[4 * o + 3] = fmaf(h, v.w, out[4 * o + 3]);
        }
    }
}

// rgb = sigmoid(x)*(1 + 2*0.001) - 0.001   (triplane.py:188,219,269)
__device__ __forceinline__ float rgb_activation(float x) { return sigmoid_fast(x) * 1.002f - 0.001f; }

// Decoder kinds as compile-time traits: padded output widths of net A / net B and which plane
// set feeds them.
template <int KIND> struct DecoderTraits;
template <> struct DecoderTraits<NFE_DEC_OSG> {            // net: 32->64->33 on the (single) plane set
    static constexpr int OUT_A = 36, OUT_B = 4, SETS = 1; static constexpr bool HAS_B = false;
};
template <> struct DecoderTraits<NFE_DEC_DISENTANGLED> {   // geo_net 32->64->16 (norm), app_net 32->64->32 (denorm)
    static constexpr int OUT_A = 16, OUT_B = 32, SETS = 2; static constexpr bool HAS_B = true;
};
template <> struct DecoderTraits<NFE_DEC_SEGMENTATION> {   // net 32->64->33 and seg_net 32->64->15, both on denorm
    static constexpr int OUT_A = 36, OUT_B = 16, SETS = 1; static constexpr bool HAS_B = true;
};

// Bilinear tap sets of the three planes for one point (q = (2/box_warp) * x).
struct Taps3 { Taps t[3]; };

__device__ __forceinline__ Taps3 taps3(float qx, float qy, float qz, int H, int W)
{
    Taps3 r;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        float gx, gy;
        project(qx, qy, qz, p, gx, gy);
        r.t[p] = plane_taps(gx, gy, H, W);
    }
    return r;
}

// The 12 taps (3 planes x 4) of one point as branch-free loads: texel offsets in float4 units,
// clamped into the plane, and weights zeroed for taps that fall outside (padding_mode='zeros').
// Branch-free matters: a conditional load per tap serialises load -> use -> load and leaves the
// gather latency-bound (measured: 2-3 loads in flight per warp); unconditional loads can all be
// issued back to back.
struct TapSet { int off4[12]; float w[12]; };

__device__ __forceinline__ TapSet make_tapset(const Taps3& tp, int H, int W)
{
    TapSet ts;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = tp.t[p].x0 + (k & 1), y = tp.t[p].y0 + (k >> 1);
            const bool inb = x >= 0 && x < W && y >= 0 && y < H;
            const int xc = min(max(x, 0), W - 1), yc = min(max(y, 0), H - 1);
            ts.off4[p * 4 + k] = ((p * H + yc) * W + xc) * (FEAT / 4);
            ts.w[p * 4 + k] = inb ? tp.t[p].w[k] : 0.0f;
        }
    }
    return ts;
}

// One point's texels from one channel-last plane set ([3,H,W,32] floats): the 8 lanes of a group
// own 4 channels each, so a warp-wide LDG.128 fetches 4 whole 128-byte texels (one per lane group).
__device__ __forceinline__ void gather_load(const float* __restrict__ set, const TapSet& ts, int c4, float4 (&v)[12])
{
    const float4* base = reinterpret_cast<const float4*>(set) + c4;
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = __ldg(base + ts.off4[i]);
}

// Bilinear blend per plane, then the plane MEAN (decoders' `.mean(1)`, triplane.py:180,211,251-252)
// for this lane's 4 channels.
__device__ __forceinline__ float4 gather_reduce(const float4 (&v)[12], const TapSet& ts)
{
    float4 f[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float w = ts.w[p * 4 + k];
            acc.x = fmaf(v[p * 4 + k].x, w, acc.x); acc.y = fmaf(v[p * 4 + k].y, w, acc.y);
            acc.z = fmaf(v[p * 4 + k].z, w, acc.z); acc.w = fmaf(v[p * 4 + k].w, w, acc.w);
        }
        f[p] = acc;
    }
    constexpr float third = 1.0f / 3.0f;
    return make_float4(((f[0].x + f[1].x) + f[2].x) * third, ((f[0].y + f[1].y) + f[2].y) * third,
                       ((f[0].z + f[1].z) + f[2].z) * third, ((f[0].w + f[1].w) + f[2].w) * third);
}

// Per-plane blends kept apart (no mean), plus each plane's sum of in-bounds tap weights: what the single-gather
// identity needs.  With zero padding, sampling an affinely transformed plane P' = s*P + m (per channel) gives
//   sample(P') = s*sample(P) + m*w_in,   w_in = sum of the in-bounds bilinear weights        
// so the de-normalised features follow from the normalised gather and the statistics alone.
__device__ __forceinline__ void gather_reduce_planes(const float4 (&v)[12], const TapSet& ts, float4 (&f)[3], float (&w_in)[3])
{
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float ws = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float w = ts.w[p * 4 + k];
            acc.x = fmaf(v[p * 4 + k].x, w, acc.x); acc.y = fmaf(v[p * 4 + k].y, w, acc.y);
            acc.z = fmaf(v[p * 4 + k].z, w, acc.z); acc.w = fmaf(v[p * 4 + k].w, w, acc.w);
            ws += w;
        }
        f[p] = acc;
        w_in[p] = ws;
    }
}

__device__ __forceinline__ float4 gather_set(const float* __restrict__ set, const TapSet& ts, int c4)
{
    float4 v[12];
    gather_load(set, ts, c4, v);
    return gather_reduce(v, ts);
}

// Per-warp staging tile: 32 samples x 32 features as float4 chunks, chunk index XOR-swizzled
// with the row so that both the gather's writes (8 lanes = one row) and the MLP's reads (one
// row per lane) are bank-conflict free.
__device__ __forceinline__ int tile_chunk(int row, int chunk) { return row * 8 + (chunk ^ (row & 7)); }

constexpr int OUT_STRIDE = 49;  // output tile row stride (floats): 1 sigma + 15/16 seg + 32 rgb, odd -> conflict free

}  // namespace nfe
