// Training-side consumers of the rendered maps: training/loss.py:28-157,276-293.
//
//   remap_seg                     BiSeNet's 19 labels -> the generator's 15 (loss.py:28-53)
//   segmentation cross-entropy    torch.nn.CrossEntropyLoss()(image_seg [N,15,H,W], labels [N,H,W]) (loss.py:276-277), fwd + bwd
//   RGB-uv histogram distances    RGBuvHistBlock ('inverse-quadratic', intensity scale) of the pixels under each semantic label's
//                                 argmax mask (or of the whole image), normalised, Hellinger distance of items 1.. to item 0
//                                 (loss.py:57-157, 284-293), fwd + bwd
//
// The reference runs the histogram loss as 12 labels x batch Python iterations of masked_select + log / abs / div + a
// [3,64,N]x[3,N,64] bmm each (launch-bound: ~40 launches per (label, item)).  Here ONE kernel builds all (label, item, channel)
// histograms: a CTA owns one 64x64 histogram, walks the image once, compacts the pixels of its label into chunks of 64 in
// shared memory (kernel values k_u, k_v per pixel and bin) and accumulates the chunk as a 64x64x64 outer-product sum in
// registers (16 cells per thread) — the same contraction the reference's bmm does, without materialising the [N,3,64] factors.
// The backward kernel has the same shape: per chunk it contracts the histogram gradient with k_v (resp. k_u) to get the
// per-pixel terms.
#include "nfe_common.cuh"

namespace nfe {

constexpr int HB = 64;                 // histogram bins per axis (RGBuvHistBlock h=64)
constexpr int HCHUNK = 64;             // pixels per accumulation step
constexpr int HTHREADS = 256;
constexpr float HEPS = 1e-6f;          // RGBuvHistBlock.EPS

__constant__ int c_seg_mapping[19] = {0, 1, 2, 2, 3, 3, 4, 5, 5, 6, 7, 8, 9, 9, 10, 11, 12, 13, 14};   // loss.py:28-48

__global__ void remap_seg_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t v = in[i];
        out[i] = (v >= 0 && v < 19) ? (int64_t)c_seg_mapping[v] : v;     // labels outside 0..18 pass through, as in the reference's loop
    }
}

// ---------------------------------------------------------------------------------------------- cross-entropy
// one thread per pixel: log-sum-exp over the C logits (stride hw), minus the picked logit
__global__ void seg_ce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int n, int c, int64_t hw, double* __restrict__ acc)
{
    const int64_t total = (int64_t)n * hw;
    double local = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / hw, p = i % hw;
        const float* x = logits + (b * c) * hw + p;
        float m = -INFINITY;
        for (int k = 0; k < c; ++k) m = fmaxf(m, __ldg(x + k * hw));
        float sum = 0.0f;
        for (int k = 0; k < c; ++k) sum += expf(__ldg(x + k * hw) - m);
        const int64_t lab = labels[i];
        const float picked = (lab >= 0 && lab < c) ? __ldg(x + lab * hw) : 0.0f;
        local += (double)((logf(sum) + m) - picked);
    }
    // block reduction, one atomic per block
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(acc, v);
    }
}

__global__ void seg_ce_finish_kernel(const double* __restrict__ acc, double count, float* __restrict__ loss)
{
    loss[0] = (float)(acc[0] / count);
}

__global__ void seg_ce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int n, int c, int64_t hw,
                                  const float* __restrict__ g_loss, float* __restrict__ g_logits)
{
    const int64_t total = (int64_t)n * hw;
    const float scale = __ldg(g_loss) / (float)total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / hw, p = i % hw;
        const float* x = logits + (b * c) * hw + p;
        float* g = g_logits + (b * c) * hw + p;
        float m = -INFINITY;
        for (int k = 0; k < c; ++k) m = fmaxf(m, __ldg(x + k * hw));
        float sum = 0.0f;
        for (int k = 0; k < c; ++k) sum += expf(__ldg(x + k * hw) - m);
        const float inv = 1.0f / sum;
        const int64_t lab = labels[i];
        for (int k = 0; k < c; ++k) g[k * hw] = (expf(__ldg(x + k * hw) - m) * inv - (k == lab ? 1.0f : 0.0f)) * scale;
    }
}

// ---------------------------------------------------------------------------------------------- RGB-uv histograms
// per-pixel quantities of RGBuvHistBlock.forward (loss.py:96-114) for colour channel c of a pixel with colours x (already in [0,1])
struct HistPixel { float iy, iu, iv; };

__device__ __forceinline__ float to_unit(float v) { return fminf(fmaxf(v / 2.0f + 0.5f, 0.0f), 1.0f); }     // loss.py:97

__device__ __forceinline__ HistPixel hist_pixel(float x0, float x1, float x2, int c)
{
    HistPixel h;
    h.iy = sqrtf(((x0 * x0 + x1 * x1) + x2 * x2) + HEPS);
    const float l0 = logf(x0 + HEPS), l1 = logf(x1 + HEPS), l2 = logf(x2 + HEPS);
    const float lc = c == 0 ? l0 : (c == 1 ? l1 : l2);
    h.iu = lc - (c == 0 ? l1 : l0);          // I[:, [1, 0, 0]]
    h.iv = lc - (c == 2 ? l1 : l2);          // I[:, [2, 2, 1]]
    return h;
}

__device__ __forceinline__ float inv_quadratic(float i, float lin, float inv_s2)
{
    const float d = fabsf(i - lin);
    return 1.0f / (1.0f + d * d * inv_s2);
}

// label of a pixel = first maximum over the seg logits (torch.argmax), or `whole` mode: every pixel belongs to the one "label"
__device__ __forceinline__ int pixel_label(const float* __restrict__ seg, int c_seg, int64_t p_total, int64_t p)
{
    int best = 0;
    float bv = __ldg(seg + p);
    for (int k = 1; k < c_seg; ++k) {
        const float v = __ldg(seg + k * p_total + p);
        if (v > bv) { bv = v; best = k; }
    }
    return best;
}

// Ordered compaction of the next (up to) HCHUNK pixels carrying this CTA's label, starting at `pos`: scans HTHREADS pixels per
// step in pixel order (deterministic), returns the number selected (sel[0..n)) and advances pos past the last pixel consumed.
struct Compactor {
    int64_t sel[HCHUNK];
    unsigned ballots[HTHREADS / 32];
    int base_of[HTHREADS / 32];
    int n_sel, stop;
};

__device__ __forceinline__ int compact_chunk(Compactor& cp, const float* __restrict__ seg, int c_seg, int64_t p_total, int label, int64_t& pos)
{
    if (threadIdx.x == 0) cp.n_sel = 0;
    __syncthreads();
    while (pos < p_total) {
        const int64_t p = pos + threadIdx.x;
        const bool mine = p < p_total && (!seg || pixel_label(seg, c_seg, p_total, p) == label);
        const unsigned bal = __ballot_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0) cp.ballots[threadIdx.x >> 5] = bal;
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = cp.n_sel, stop = HTHREADS;
            for (int w = 0; w < HTHREADS / 32 && stop == HTHREADS; ++w) {
                cp.base_of[w] = n;
                const int cnt = __popc(cp.ballots[w]);
                if (n + cnt >= HCHUNK) {
                    // the chunk fills inside (or right at the end of) this warp: stop after the pixel that fills it
                    unsigned m = cp.ballots[w];
                    int room = HCHUNK - n, lane = -1;
                    while (room > 0) { lane = __ffs(m) - 1; m &= m - 1; --room; }
                    stop = w * 32 + lane + 1;
                    n = HCHUNK;
                } else {
                    n += cnt;
                }
            }
            cp.n_sel = n;
            cp.stop = stop;
        }
        __syncthreads();
        const int stop = cp.stop;
        if (mine && (int)threadIdx.x < stop) cp.sel[cp.base_of[threadIdx.x >> 5] + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = p;
        pos += stop;
        __syncthreads();
        if (cp.n_sel == HCHUNK) break;
    }
    return cp.n_sel;
}

struct HistArgs {
    const float* img;        // [B,3,P] in (-1,1)
    const float* seg;        // [B,C,P] logits, or NULL (whole image: n_labels == 1)
    const int* label_ids;    // [n_labels] device
    const float* lin;        // [64] torch.linspace(-3, 3, 64)
    int b, c_seg, n_labels;
    int64_t p;
    float inv_s2;            // 1 / sigma^2
};

// grid = (3 channels, B items, n_labels); hist_raw [n_labels, B, 3, 64, 64]
__global__ void __launch_bounds__(HTHREADS) hist_fwd_kernel(HistArgs a, float* __restrict__ hist_raw)
{
    __shared__ float ku[HCHUNK][HB + 1], kv[HCHUNK][HB + 1];
    __shared__ Compactor cp;
    const int c = blockIdx.x, item = blockIdx.y, li = blockIdx.z;
    const int label = a.seg ? __ldg(a.label_ids + li) : 0;
    const float* img = a.img + (int64_t)item * 3 * a.p;
    const float* seg = a.seg ? a.seg + (int64_t)item * a.c_seg * a.p : nullptr;
    const int tu = threadIdx.x >> 4, tv = threadIdx.x & 15;          // thread owns cells u = tu + 16*i, v = tv + 16*j (4 x 4)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    int64_t pos = 0;
    while (pos < a.p) {
        const int n_pix = compact_chunk(cp, seg, a.c_seg, a.p, label, pos);
        if (n_pix == 0) break;
        // ---- kernel values of the chunk: thread t fills (pixel t/4, bins (t%4)*16 .. +15)
        {
            const int px = threadIdx.x >> 2, b0 = (threadIdx.x & 3) * 16;
            if (px < n_pix) {
                const int64_t p = cp.sel[px];
                const HistPixel h = hist_pixel(to_unit(__ldg(img + p)), to_unit(__ldg(img + a.p + p)), to_unit(__ldg(img + 2 * a.p + p)), c);
                for (int k = 0; k < 16; ++k) {
                    const float lin = __ldg(a.lin + b0 + k);
                    ku[px][b0 + k] = h.iy * inv_quadratic(h.iu, lin, a.inv_s2);
                    kv[px][b0 + k] = inv_quadratic(h.iv, lin, a.inv_s2);
                }
            } else {
                for (int k = 0; k < 16; ++k) { ku[px][b0 + k] = 0.0f; kv[px][b0 + k] = 0.0f; }
            }
        }
        __syncthreads();
        // ---- 64 x 64 x (chunk) outer-product accumulation
        for (int px = 0; px < n_pix; ++px) {
            float u4[4], v4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { u4[i] = ku[px][tu + 16 * i]; v4[i] = kv[px][tv + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(u4[i], v4[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = hist_raw + (((int64_t)li * a.b + item) * 3 + c) * HB * HB;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[(tu + 16 * i) * HB + tv + 16 * j] = acc[i][j];
}

// normalise each (label, item) histogram (3 x 64 x 64 cells) by its sum + EPS; grid = n_labels * B
__global__ void hist_normalize_kernel(const float* __restrict__ raw, float* __restrict__ norm, float* __restrict__ totals)
{
    const int64_t base = (int64_t)blockIdx.x * 3 * HB * HB;
    __shared__ double red[32];
    double local = 0.0;
    for (int i = threadIdx.x; i < 3 * HB * HB; i += blockDim.x) local += (double)raw[base + i];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const float tot = (float)red[0];
    if (threadIdx.x == 0) totals[blockIdx.x] = tot;
    const float inv = 1.0f / (tot + HEPS);
    for (int i = threadIdx.x; i < 3 * HB * HB; i += blockDim.x) norm[base + i] = raw[base + i] * inv;
}

// Hellinger distance of items 1..B-1 to item 0, per label (loss.py:123-126): dist[l] = (1/sqrt 2) sqrt(S_l) / (B-1); grid = n_labels
__global__ void hist_dist_kernel(const float* __restrict__ norm, int b, float* __restrict__ s_out, float* __restrict__ dist)
{
    const int64_t cells = 3 * HB * HB;
    const float* h0 = norm + (int64_t)blockIdx.x * b * cells;
    __shared__ double red[32];
    double local = 0.0;
    for (int64_t i = threadIdx.x; i < (int64_t)(b - 1) * cells; i += blockDim.x) {
        const float t = sqrtf(h0[i % cells]), v = sqrtf(h0[cells + i]);
        const float d = t - v;
        local += (double)(d * d);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) {
            s_out[blockIdx.x] = (float)v;
            dist[blockIdx.x] = (float)(0.70710678118654752440 * sqrt(v) / (double)(b - 1));     // b == 1: 0/0 = NaN, as in the reference
        }
    }
}

// loss = sum_l w_l * dist_l (compute_seg_hist_dist's label loop, loss.py:142-153), in label order like the reference's Python sum
__global__ void hist_weighted_sum_kernel(const float* __restrict__ dist, const float* __restrict__ weights, int n_labels, float* __restrict__ loss)
{
    float acc = 0.0f;
    for (int l = 0; l < n_labels; ++l) acc = acc + weights[l] * dist[l];
    loss[0] = acc;
}

// d loss / d raw histogram of items 1.. (item 0 is the detached target): through the distance and the normalisation.
//   loss = sum_l w_l * g_loss * dist_l;   d dist / d hn = k / ((B-1) 2 sqrt(S)) * (1 - sqrt(t) / sqrt(hn))
//   hn = h / (T + eps):  d / d h_i = g_i / (T+eps) - sum_k g_k h_k / (T+eps)^2
// grid = n_labels * B
__global__ void hist_dist_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ norm, const float* __restrict__ totals,
                                     const float* __restrict__ s_in, const float* __restrict__ weights, const float* __restrict__ g_loss, int b,
                                     float* __restrict__ g_raw)
{
    const int li = blockIdx.x / b, item = blockIdx.x % b;
    const int64_t cells = 3 * HB * HB, base = (int64_t)blockIdx.x * cells;
    if (item == 0) {
        for (int i = threadIdx.x; i < cells; i += blockDim.x) g_raw[base + i] = 0.0f;
        return;
    }
    const float* h0 = norm + (int64_t)li * b * cells;
    const float S = s_in[li];
    const float coef = S > 0.0f ? __ldg(weights + li) * __ldg(g_loss) * 0.70710678118654752440f / ((float)(b - 1) * 2.0f * sqrtf(S)) : 0.0f;
    const float tot = totals[blockIdx.x] + HEPS;
    __shared__ double red[32];
    double local = 0.0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const float hn = norm[base + i];
        const float g = hn > 0.0f ? coef * (1.0f - sqrtf(h0[i]) / sqrtf(hn)) : 0.0f;
        local += (double)(g * raw[base + i]);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const float dot = (float)red[0];
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const float hn = norm[base + i];
        const float g = hn > 0.0f ? coef * (1.0f - sqrtf(h0[i]) / sqrtf(hn)) : 0.0f;
        g_raw[base + i] = g / tot - dot / (tot * tot);
    }
}

// d loss / d image from d loss / d raw histograms.  Same walk as the forward: a CTA owns (channel c, item, label), compacts its
// pixels into chunks of 64 and, per pixel, contracts G[c] (64 x 64, in shared memory) with the kernel vectors:
//   T_u = sum_v G[u][v] k_v[v]     S  = sum_u k_u[u] T_u      (d / d Iy)       A = Iy sum_u k_u'[u] T_u   (d / d Iu)
//   R_v = sum_u G[u][v] k_u[u]     B = Iy sum_v k_v'[v] R_v   (d / d Iv)       with k' = -2 (I - lin) / sigma^2 * k^2
// and chains through Iy, Iu, Iv and the clamp to the three colours of the pixel (atomicAdd: the three channel CTAs of an
// (item, label) all write the pixel).  grid = (3, B, n_labels); item 0 returns at once.
__global__ void __launch_bounds__(HTHREADS) hist_bwd_kernel(HistArgs a, const float* __restrict__ g_raw, float* __restrict__ g_img)
{
    const int c = blockIdx.x, item = blockIdx.y, li = blockIdx.z;
    if (item == 0) return;
    extern __shared__ float hsm[];
    float (*G)[HB + 1] = reinterpret_cast<float (*)[HB + 1]>(hsm);                 // [64][65] gradient of this CTA's histogram
    float (*ku)[HB + 1] = G + HB;                                                  // [64 pixels][65] k_u (without Iy)
    float (*kv)[HB + 1] = ku + HCHUNK;                                             // [64 pixels][65] k_v
    __shared__ Compactor cp;
    const int label = a.seg ? __ldg(a.label_ids + li) : 0;
    const float* img = a.img + (int64_t)item * 3 * a.p;
    float* gim = g_img + (int64_t)item * 3 * a.p;
    const float* seg = a.seg ? a.seg + (int64_t)item * a.c_seg * a.p : nullptr;
    const float* g = g_raw + (((int64_t)li * a.b + item) * 3 + c) * HB * HB;
    for (int i = threadIdx.x; i < HB * HB; i += blockDim.x) G[i / HB][i % HB] = g[i];
    __syncthreads();
    int64_t pos = 0;
    while (pos < a.p) {
        const int n_pix = compact_chunk(cp, seg, a.c_seg, a.p, label, pos);
        if (n_pix == 0) break;
        // ---- 4 threads per pixel: thread q of a pixel handles bins q, q+4, ... (16 of them)
        const int px = threadIdx.x >> 2, q = threadIdx.x & 3;
        float s_part = 0.0f, a_part = 0.0f, b_part = 0.0f;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        HistPixel h = {1.0f, 0.0f, 0.0f};
        int64_t p = 0;
        if (px < n_pix) {
            p = cp.sel[px];
            x0 = to_unit(__ldg(img + p)); x1 = to_unit(__ldg(img + a.p + p)); x2 = to_unit(__ldg(img + 2 * a.p + p));
            h = hist_pixel(x0, x1, x2, c);
            for (int k = 0; k < 16; ++k) {
                const float lin = __ldg(a.lin + q + 4 * k);
                ku[px][q + 4 * k] = inv_quadratic(h.iu, lin, a.inv_s2);
                kv[px][q + 4 * k] = inv_quadratic(h.iv, lin, a.inv_s2);
            }
        }
        __syncwarp();                                       // the 4 threads of a pixel sit in one warp
        if (px < n_pix) {
            for (int k = 0; k < 16; ++k) {
                const int u = q + 4 * k;
                const float lin_u = __ldg(a.lin + u);
                float t = 0.0f, r = 0.0f;                  // T_u: row u of G against k_v;  R_u: column u of G against k_u
                for (int w = 0; w < HB; ++w) {
                    t = fmaf(G[u][w], kv[px][w], t);
                    r = fmaf(G[w][u], ku[px][w], r);
                }
                const float kuu = ku[px][u], kvu = kv[px][u];
                s_part = fmaf(kuu, t, s_part);
                a_part = fmaf(-2.0f * (h.iu - lin_u) * a.inv_s2 * kuu * kuu, t, a_part);
                b_part = fmaf(-2.0f * (h.iv - lin_u) * a.inv_s2 * kvu * kvu, r, b_part);
            }
        }
        // reduce over the 4 threads of the pixel
        for (int o = 1; o < 4; o <<= 1) {
            s_part += __shfl_xor_sync(0xffffffffu, s_part, o);
            a_part += __shfl_xor_sync(0xffffffffu, a_part, o);
            b_part += __shfl_xor_sync(0xffffffffu, b_part, o);
        }
        if (px < n_pix && q == 0) {
            const float d_iy = s_part, d_iu = h.iy * a_part, d_iv = h.iy * b_part;
            // Iy = sqrt(sum x^2 + eps);  Iu = log(x_c + eps) - log(x_cu + eps), cu = {1,0,0}[c];  Iv likewise with cv = {2,2,1}[c]
            const float xs[3] = {x0, x1, x2};
            float gx[3] = {d_iy * x0 / h.iy, d_iy * x1 / h.iy, d_iy * x2 / h.iy};
            const int cu = c == 0 ? 1 : 0, cv = c == 2 ? 1 : 2;
            gx[c] += (d_iu + d_iv) / (xs[c] + HEPS);
            gx[cu] -= d_iu / (xs[cu] + HEPS);
            gx[cv] -= d_iv / (xs[cv] + HEPS);
            // x = clamp(v/2 + 0.5, 0, 1): slope 1/2 inside, 0 where clamped (torch.clamp passes gradient on the closed interval)
            for (int k = 0; k < 3; ++k) {
                const float v = __ldg(img + k * a.p + p) / 2.0f + 0.5f;
                if (v >= 0.0f && v <= 1.0f) atomicAdd(gim + k * a.p + p, 0.5f * gx[k]);
            }
        }
        __syncthreads();
    }
}

}  // namespace nfe

using namespace nfe;

NFE_EXPORT int nfe_remap_seg(const int64_t* labels19, int64_t n, int64_t* out, nfe_stream_t stream)
{
    if (n == 0) return 0;
    NFE_REQUIRE(labels19 && out && n > 0, "nfe_remap_seg: bad arguments");
    remap_seg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(labels19, n, out);
    NFE_LAUNCH_CHECK("remap_seg_kernel");
    return 0;
}

NFE_EXPORT int nfe_seg_cross_entropy_fwd(const float* logits, const int64_t* labels, int n, int c, int64_t hw, float* loss, double* acc_ws,
                                         nfe_stream_t stream)
{
    NFE_REQUIRE(logits && labels && loss && acc_ws && n > 0 && c > 0 && hw > 0, "nfe_seg_cross_entropy_fwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(acc_ws, 0, sizeof(double), st);
    const int64_t total = (int64_t)n * hw;
    const unsigned grid = (unsigned)min((int64_t)sm_count() * 8, (total + 255) / 256);
    seg_ce_fwd_kernel<<<grid, 256, 0, st>>>(logits, labels, n, c, hw, acc_ws);
    NFE_LAUNCH_CHECK("seg_ce_fwd_kernel");
    seg_ce_finish_kernel<<<1, 1, 0, st>>>(acc_ws, (double)total, loss);
    NFE_LAUNCH_CHECK("seg_ce_finish_kernel");
    return 0;
}

NFE_EXPORT int nfe_seg_cross_entropy_bwd(const float* logits, const int64_t* labels, int n, int c, int64_t hw, const float* g_loss,
                                         float* g_logits, nfe_stream_t stream)
{
    NFE_REQUIRE(logits && labels && g_loss && g_logits && n > 0 && c > 0 && hw > 0, "nfe_seg_cross_entropy_bwd: bad arguments");
    const int64_t total = (int64_t)n * hw;
    const unsigned grid = (unsigned)min((int64_t)sm_count() * 8, (total + 255) / 256);
    seg_ce_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, labels, n, c, hw, g_loss, g_logits);
    NFE_LAUNCH_CHECK("seg_ce_bwd_kernel");
    return 0;
}

static int hist_args(HistArgs& a, const char* who, const float* img, const float* seg, const int* label_ids, const float* lin, int b, int c_seg,
                     int n_labels, int64_t p, float sigma)
{
    NFE_REQUIRE(img && lin && b >= 1 && p >= 1 && n_labels >= 1 && sigma > 0.0f, "%s: bad arguments", who);
    NFE_REQUIRE(seg ? (label_ids && c_seg >= 1) : n_labels == 1, "%s: per-label histograms need seg logits and label ids; whole-image mode has one label", who);
    a.img = img; a.seg = seg; a.label_ids = label_ids; a.lin = lin; a.b = b; a.c_seg = c_seg; a.n_labels = n_labels; a.p = p;
    a.inv_s2 = 1.0f / (sigma * sigma);
    return 0;
}

NFE_EXPORT int nfe_hist_dist_fwd(const float* img, const float* seg, const int* label_ids, const float* lin, int b, int c_seg, int n_labels,
                                 int64_t p, float sigma, const float* weights, float* hist_raw, float* hist_norm, float* totals, float* s_ws,
                                 float* dist, float* loss, nfe_stream_t stream)
{
    HistArgs a;
    if (int rc = hist_args(a, "nfe_hist_dist_fwd", img, seg, label_ids, lin, b, c_seg, n_labels, p, sigma)) return rc;
    NFE_REQUIRE(weights && hist_raw && hist_norm && totals && s_ws && dist && loss, "nfe_hist_dist_fwd: null pointer");
    cudaStream_t st = as_stream(stream);
    hist_fwd_kernel<<<dim3(3, b, n_labels), HTHREADS, 0, st>>>(a, hist_raw);
    NFE_LAUNCH_CHECK("hist_fwd_kernel");
    hist_normalize_kernel<<<n_labels * b, 256, 0, st>>>(hist_raw, hist_norm, totals);
    NFE_LAUNCH_CHECK("hist_normalize_kernel");
    hist_dist_kernel<<<n_labels, 256, 0, st>>>(hist_norm, b, s_ws, dist);
    NFE_LAUNCH_CHECK("hist_dist_kernel");
    hist_weighted_sum_kernel<<<1, 1, 0, st>>>(dist, weights, n_labels, loss);
    NFE_LAUNCH_CHECK("hist_weighted_sum_kernel");
    return 0;
}

NFE_EXPORT int nfe_hist_dist_bwd(const float* img, const float* seg, const int* label_ids, const float* lin, int b, int c_seg, int n_labels,
                                 int64_t p, float sigma, const float* hist_raw, const float* hist_norm, const float* totals, const float* s_ws,
                                 const float* weights, const float* g_loss, float* g_raw_ws, float* g_img, nfe_stream_t stream)
{
    HistArgs a;
    if (int rc = hist_args(a, "nfe_hist_dist_bwd", img, seg, label_ids, lin, b, c_seg, n_labels, p, sigma)) return rc;
    NFE_REQUIRE(hist_raw && hist_norm && totals && s_ws && weights && g_loss && g_raw_ws && g_img, "nfe_hist_dist_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
    hist_dist_bwd_kernel<<<n_labels * b, 256, 0, st>>>(hist_raw, hist_norm, totals, s_ws, weights, g_loss, b, g_raw_ws);
    NFE_LAUNCH_CHECK("hist_dist_bwd_kernel");
    const size_t smem = sizeof(float) * (HB + 2 * HCHUNK) * (HB + 1);           // 49.9 KB: above the 48 KB default
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(hist_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        NFE_REQUIRE(e == cudaSuccess, "nfe_hist_dist_bwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    hist_bwd_kernel<<<dim3(3, b, n_labels), HTHREADS, smem, st>>>(a, g_raw_ws, g_img);      // g_img zero-initialised by the caller
    NFE_LAUNCH_CHECK("hist_bwd_kernel");
    return 0;
}
