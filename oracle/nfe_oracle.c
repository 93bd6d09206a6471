/*
 * nfe_oracle.c — CPU restatement of NeRFFaceEditing's tri-plane volume-rendering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the timed CPU baseline.  The product path
 * (nerffaceediting_b200/) never imports it and has no CPU fallback.
 *
 * Parity status: PINNED against outputs of the reference itself.  The reference ships no
 * tests or golden vectors for this path, so tests/golden/make_golden.py
 * imports the unmodified reference from /root/reference, runs it on seeded inputs (with
 * deterministic sampling imposed from outside) and commits the inputs/outputs as fixtures;
 * tests/test_oracle_golden.py checks every function below against them.
 *
 * All arithmetic is fp32 unless a comment says otherwise; the file is compiled with
 * -ffp-contract=off so that no multiply-add is fused behind the source's back.  Each
 * function cites the reference file:line it restates (paths relative to the reference root).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NFO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Plane statistics — training/triplane.py:56-68 (twins utils.py:146-158)
 *   mean over H*W; "var" is sqrt of the UNBIASED variance (divisor HW-1);
 *   norm = (x - mean) / (std + 1e-8);  denorm = norm * std' + mean'.
 * Accumulation is in double and rounded once (torch's fp32 cascade sum agrees to ~1e-7).
 * ---------------------------------------------------------------------------------------- */
NFO_API void nfo_plane_stats(const float* planes, int64_t n_slabs, int64_t hw, float* mean, float* std_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n_slabs; ++s) {
        const float* x = planes + s * hw;
        double sum = 0.0;
        for (int64_t i = 0; i < hw; ++i) sum += (double)x[i];
        const double m = sum / (double)hw;
        double ss = 0.0;
        for (int64_t i = 0; i < hw; ++i) { const double d = (double)x[i] - m; ss += d * d; }
        mean[s] = (float)m;
        std_out[s] = sqrtf((float)(ss / (double)(hw - 1)));
    }
}

NFO_API void nfo_normalize(const float* planes, const float* mean, const float* std_in,
                           int64_t n_slabs, int64_t hw, float* out)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n_slabs; ++s) {
        const float m = mean[s];
        const float d = std_in[s] + 1e-8f;
        for (int64_t i = 0; i < hw; ++i) out[s * hw + i] = (planes[s * hw + i] - m) / d;
    }
}

/* stat_slabs: number of (n,c) entries in mean/std; slab s uses entry s % stat_slabs, which
 * covers both the per-item case (stat_slabs == n_slabs) and the broadcast of one item's
 * statistics over the batch (triplane.py:100-101, stat_slabs == channels). */
NFO_API void nfo_denormalize(const float* norm, const float* mean, const float* std_in,
                             int64_t n_slabs, int64_t stat_slabs, int64_t hw, float* out)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n_slabs; ++s) {
        const float m = mean[s % stat_slabs];
        const float d = std_in[s % stat_slabs];
        for (int64_t i = 0; i < hw; ++i) {
            const float t = norm[s * hw + i] * d;
            out[s * hw + i] = t + m;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Ray generation — training/volumetric_rendering/ray_sampler.py:24-63
 *   ray m = i*res + j  (i row, j column); x_cam = (j+.5)/res, y_cam = (i+.5)/res, z = 1.
 * ---------------------------------------------------------------------------------------- */
NFO_API void nfo_generate_rays(const float* cam2world /*[N,4,4]*/, const float* intrinsics /*[N,3,3]*/,
                               int n, int res, float* origins /*[N,res*res,3]*/, float* dirs)
{
    const float inv_res = 1.0f / (float)res;
    const float half = 0.5f / (float)res;
    for (int b = 0; b < n; ++b) {
        const float* c = cam2world + 16 * b;
        const float* k = intrinsics + 9 * b;
        const float fx = k[0], sk = k[1], cx = k[2], fy = k[4], cy = k[5];
        for (int i = 0; i < res; ++i) {
            for (int j = 0; j < res; ++j) {
                const float x_cam = (float)j * inv_res + half;
                const float y_cam = (float)i * inv_res + half;
                /* ray_sampler.py:51-52 */
                const float x_lift = (x_cam - cx + cy * sk / fy - sk * y_cam / fy) / fx * 1.0f;
                const float y_lift = (y_cam - cy) / fy * 1.0f;
                const float p[4] = { x_lift, y_lift, 1.0f, 1.0f };
                float w[3];
                for (int r = 0; r < 3; ++r) {
                    float acc = 0.0f;
                    for (int q = 0; q < 4; ++q) acc += c[4 * r + q] * p[q];
                    w[r] = acc;
                }
                float d[3] = { w[0] - c[3], w[1] - c[7], w[2] - c[11] };
                float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                if (nrm < 1e-12f) nrm = 1e-12f;     /* F.normalize eps */
                const int64_t m = ((int64_t)b * res * res + (int64_t)i * res + j) * 3;
                for (int r = 0; r < 3; ++r) { dirs[m + r] = d[r] / nrm; }
                origins[m + 0] = c[3]; origins[m + 1] = c[7]; origins[m + 2] = c[11];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Ray / axis-aligned-box limits — training/volumetric_rendering/math_utils.py:46-98
 *   slab test against [-L/2, L/2]^3; invalid rays get (tmin,tmax) = (-1,-2).
 * ---------------------------------------------------------------------------------------- */
NFO_API void nfo_ray_limits_box(const float* origins, const float* dirs, int64_t n_rays,
                                float box_side_length, float* tmin_out, float* tmax_out)
{
    const float lo = -1.0f * (box_side_length / 2.0f), hi = 1.0f * (box_side_length / 2.0f);
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = origins + 3 * r;
        const float* d = dirs + 3 * r;
        float inv[3]; int sg[3];
        for (int a = 0; a < 3; ++a) { inv[a] = 1.0f / d[a]; sg[a] = inv[a] < 0.0f; }
        int valid = 1;
        float tmin = ((sg[0] ? hi : lo) - o[0]) * inv[0];
        float tmax = ((sg[0] ? lo : hi) - o[0]) * inv[0];
        const float tymin = ((sg[1] ? hi : lo) - o[1]) * inv[1];
        const float tymax = ((sg[1] ? lo : hi) - o[1]) * inv[1];
        if (tmin > tymax || tymin > tmax) valid = 0;
        /* torch.max / torch.min propagate NaN; fmaxf does not, so spell it out */
        tmin = (tmin != tmin || tymin != tymin) ? NAN : (tmin > tymin ? tmin : tymin);
        tmax = (tmax != tmax || tymax != tymax) ? NAN : (tmax < tymax ? tmax : tymax);
        const float tzmin = ((sg[2] ? hi : lo) - o[2]) * inv[2];
        const float tzmax = ((sg[2] ? lo : hi) - o[2]) * inv[2];
        if (tmin > tzmax || tzmin > tmax) valid = 0;
        tmin = (tmin != tmin || tzmin != tzmin) ? NAN : (tmin > tzmin ? tmin : tzmin);
        tmax = (tmax != tmax || tzmax != tzmax) ? NAN : (tmax < tzmax ? tmax : tzmax);
        if (!valid) { tmin = -1.0f; tmax = -2.0f; }
        tmin_out[r] = tmin; tmax_out[r] = tmax;
    }
}

/* ------------------------------------------------------------------------------------------
 * Coarse depths — training/volumetric_rendering/renderer.py:169-192
 *   scalar limits : t = table[s] + u*delta            (table = torch.linspace on the host)
 *   tensor limits : t = start + (s/(S-1))*(stop-start) + u*(stop-start)/(S-1)   (math_utils.py:101-118)
 *   disparity     : s' = table01[s] + u/(S-1);  t = 1 / (1/start*(1-s') + 1/end*s')
 * jitter == NULL means the deterministic parity mode (u := 0).
 * ---------------------------------------------------------------------------------------- */
NFO_API void nfo_sample_stratified(int64_t n_rays, int s_c, int mode /*0 scalar,1 per-ray,2 disparity*/,
                                   const float* table /*[S] for modes 0,2*/,
                                   double ray_start, double ray_end,
                                   const float* start_per_ray, const float* end_per_ray,
                                   const float* jitter /*[n_rays,S] or NULL*/, float* depths /*[n_rays,S]*/)
{
    for (int64_t r = 0; r < n_rays; ++r) {
        for (int s = 0; s < s_c; ++s) {
            const float u = jitter ? jitter[r * s_c + s] : 0.0f;
            float t;
            if (mode == 0) {
                const float delta = (float)((ray_end - ray_start) / (double)(s_c - 1));
                t = table[s] + u * delta;
            } else if (mode == 1) {
                const float a = start_per_ray[r], b = end_per_ray[r];
                const float step = (float)s / (float)(s_c - 1);
                t = a + step * (b - a);
                const float delta = (b - a) / (float)(s_c - 1);
                t = t + u * delta;
            } else {
                const float delta = (float)(1.0 / (double)(s_c - 1));
                const float sp = table[s] + u * delta;
                const float ia = (float)(1.0 / ray_start), ib = (float)(1.0 / ray_end);
                t = 1.0f / (ia * (1.0f - sp) + ib * sp);
            }
            depths[r * s_c + s] = t;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Tri-plane bilinear gather — training/volumetric_rendering/renderer.py:23-65
 *   q = (2/box_warp) * x;  plane 0 -> (q.x,q.y), plane 1 -> (q.x,q.z), plane 2 -> (q.z,q.x)
 *   (generate_planes + linalg.inv + [..., :2], renderer.py:29-53); first grid channel indexes W.
 *   grid_sample(bilinear, zeros, align_corners=False): ix = ((g+1)*W - 1)/2.
 * One tap set for one sample of one plane; `stride_c/stride_y/stride_x` let the same code
 * read NCHW (reference layout) or channel-last staging buffers.
 * ---------------------------------------------------------------------------------------- */
static inline void nfo_plane_taps(float gx, float gy, int H, int W, int* x0, int* y0, float w[4])
{
    const float ix = ((gx + 1.0f) * (float)W - 1.0f) / 2.0f;
    const float iy = ((gy + 1.0f) * (float)H - 1.0f) / 2.0f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
    w[0] = (fx1 - ix) * (fy1 - iy);   /* nw : (x0,y0) */
    w[1] = (ix - fx0) * (fy1 - iy);   /* ne : (x1,y0) */
    w[2] = (fx1 - ix) * (iy - fy0);   /* sw : (x0,y1) */
    w[3] = (ix - fx0) * (iy - fy0);   /* se : (x1,y1) */
    /* clamp before the int conversion so far-away samples cannot overflow */
    const float cx = fx0 < -2.0f ? -2.0f : (fx0 > (float)W ? (float)W : fx0);
    const float cy = fy0 < -2.0f ? -2.0f : (fy0 > (float)H ? (float)H : fy0);
    *x0 = (int)cx; *y0 = (int)cy;
    if (!(ix == ix)) { *x0 = -2; }    /* NaN coordinate: every tap out of bounds */
    if (!(iy == iy)) { *y0 = -2; }
}

static inline void nfo_project(const float q[3], int plane, float* gx, float* gy)
{
    if (plane == 0) { *gx = q[0]; *gy = q[1]; }
    else if (plane == 1) { *gx = q[0]; *gy = q[2]; }
    else { *gx = q[2]; *gy = q[0]; }
}

/* features of one plane at one point: out[c] = sum of in-bounds taps */
static inline void nfo_gather_plane(const float* plane, int C, int H, int W,
                                    int64_t stride_c, int64_t stride_y, int64_t stride_x,
                                    float gx, float gy, float* out)
{
    int x0, y0; float w[4];
    nfo_plane_taps(gx, gy, H, W, &x0, &y0, w);
    for (int c = 0; c < C; ++c) out[c] = 0.0f;
    for (int t = 0; t < 4; ++t) {
        const int x = x0 + (t & 1), y = y0 + (t >> 1);
        if (x < 0 || x > W - 1 || y < 0 || y > H - 1) continue;
        const float* p = plane + (int64_t)y * stride_y + (int64_t)x * stride_x;
        for (int c = 0; c < C; ++c) out[c] += p[c * stride_c] * w[t];
    }
}

NFO_API void nfo_sample_planes(const float* planes /*[N,3,C,H,W]*/, const float* coords /*[N,M,3]*/,
                               int n, int64_t m, int C, int H, int W, float box_warp,
                               float* out /*[N,3,M,C]*/)
{
    const float scale = (float)(2.0 / (double)box_warp);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < n; ++b) {
        for (int64_t i = 0; i < m; ++i) {
            const float* x = coords + ((int64_t)b * m + i) * 3;
            const float q[3] = { scale * x[0], scale * x[1], scale * x[2] };
            for (int p = 0; p < 3; ++p) {
                float gx, gy; nfo_project(q, p, &gx, &gy);
                const float* plane = planes + ((int64_t)b * 3 + p) * C * H * W;
                nfo_gather_plane(plane, C, H, W, (int64_t)H * W, W, 1, gx, gy,
                                 out + (((int64_t)b * 3 + p) * m + i) * C);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Decoders — training/triplane.py:167-270 on top of FullyConnectedLayer
 * (training/networks_stylegan2.py:96-127):  w = weight*weight_gain, b = bias*bias_gain,
 * y = b + x w^T.  Softplus(beta=1, threshold=20);  rgb = sigmoid(.)*(1+2*0.001) - 0.001.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float* w1; const float* b1;   /* [hidden,in], [hidden] */
    const float* w2; const float* b2;   /* [out,hidden], [out]   */
    int in_dim, hidden, out_dim;
    float wgain1, bgain1, wgain2, bgain2;
} nfo_mlp;

static inline float nfo_softplus(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
static inline float nfo_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

/* y[out] = FC2(Softplus(FC1(x[in])));  gains applied to the parameters in fp32 first. */
/* Normative definition, kept for reading; the callers below use nfo_mlp_eval_fast. */
static __attribute__((unused)) void nfo_mlp_eval(const nfo_mlp* p, const float* x, float* y)
{
    float h[256];
    for (int j = 0; j < p->hidden; ++j) {
        float acc = 0.0f;
        for (int k = 0; k < p->in_dim; ++k) acc += x[k] * (p->w1[j * p->in_dim + k] * p->wgain1);
        const float b = p->bgain1 != 1.0f ? p->b1[j] * p->bgain1 : p->b1[j];
        h[j] = nfo_softplus(b + acc);
    }
    for (int o = 0; o < p->out_dim; ++o) {
        float acc = 0.0f;
        for (int j = 0; j < p->hidden; ++j) acc += h[j] * (p->w2[o * p->hidden + j] * p->wgain2);
        const float b = p->bgain2 != 1.0f ? p->b2[o] * p->bgain2 : p->b2[o];
        y[o] = b + acc;
    }
}

/* The same arithmetic, laid out for speed (the CPU-baseline legs of bench.py time this file):
 * gains folded once, weights transposed so that the loops over hidden units / outputs are the
 * contiguous (vectorisable) ones.  Every output still accumulates its products in ascending
 * k (resp. j) order starting from 0, so results are bit-identical to nfo_mlp_eval. */
typedef struct {
    float w1t[64 * 256];   /* [in][hidden]  */
    float b1[256];
    float w2t[256 * 64];   /* [hidden][out] */
    float b2[64];
    int in_dim, hidden, out_dim;
} nfo_mlp_fast;

static int nfo_mlp_prepare(const nfo_mlp* p, nfo_mlp_fast* f)
{
    if (!p) { f->in_dim = f->hidden = f->out_dim = 0; return 0; }
    if (p->in_dim > 64 || p->hidden > 256 || p->out_dim > 64) return 1;
    f->in_dim = p->in_dim; f->hidden = p->hidden; f->out_dim = p->out_dim;
    for (int j = 0; j < p->hidden; ++j) {
        for (int k = 0; k < p->in_dim; ++k) f->w1t[k * p->hidden + j] = p->w1[j * p->in_dim + k] * p->wgain1;
        f->b1[j] = p->bgain1 != 1.0f ? p->b1[j] * p->bgain1 : p->b1[j];
    }
    for (int o = 0; o < p->out_dim; ++o) {
        for (int j = 0; j < p->hidden; ++j) f->w2t[j * p->out_dim + o] = p->w2[o * p->hidden + j] * p->wgain2;
        f->b2[o] = p->bgain2 != 1.0f ? p->b2[o] * p->bgain2 : p->b2[o];
    }
    return 0;
}

static void nfo_mlp_eval_fast(const nfo_mlp_fast* f, const float* x, float* y)
{
    float h[256], acc[256];
    const int H = f->hidden, O = f->out_dim;
    for (int j = 0; j < H; ++j) acc[j] = 0.0f;
    for (int k = 0; k < f->in_dim; ++k) {
        const float xk = x[k];
        const float* w = f->w1t + k * H;
        for (int j = 0; j < H; ++j) acc[j] += xk * w[j];
    }
    for (int j = 0; j < H; ++j) h[j] = nfo_softplus(f->b1[j] + acc[j]);
    for (int o = 0; o < O; ++o) acc[o] = 0.0f;
    for (int j = 0; j < H; ++j) {
        const float hj = h[j];
        const float* w = f->w2t + j * O;
        for (int o = 0; o < O; ++o) acc[o] += hj * w[o];
    }
    for (int o = 0; o < O; ++o) y[o] = f->b2[o] + acc[o];
}

NFO_API void nfo_fc(const float* x, int64_t rows, int in_dim, int out_dim, const float* weight,
                    const float* bias, float wgain, float bgain, float* y)
{
    for (int64_t r = 0; r < rows; ++r)
        for (int o = 0; o < out_dim; ++o) {
            float acc = 0.0f;
            for (int k = 0; k < in_dim; ++k) acc += x[r * in_dim + k] * (weight[o * in_dim + k] * wgain);
            const float b = bgain != 1.0f ? bias[o] * bgain : bias[o];
            y[r * out_dim + o] = b + acc;
        }
}

enum { NFO_DEC_OSG = 0, NFO_DEC_DISENTANGLED = 1, NFO_DEC_SEGMENTATION = 2 };

/* One sample.  f_norm / f_denorm are the plane-MEANS of the gathered features (mean over the
 * 3 planes, triplane.py:180,211,251-252).  Output widths: rgb = color_dim, seg = seg_dim.
 *   OSG            (triplane.py:178-190): net(f_denorm)      -> sigma = y[0], rgb = sig(y[1:])
 *   Disentangled   (triplane.py:249-270): geo_net(f_norm)    -> sigma = g[0], seg = g[1:]
 *                                         app_net(f_denorm)  -> rgb = sig(a)
 *   Segmentation   (triplane.py:209-230): net(f_denorm) as OSG; seg_net(f_denorm) -> seg
 * For OSG the single feature tensor is passed as f_denorm. */
static inline void nfo_decode_sample(int kind, const nfo_mlp_fast* net_a, const nfo_mlp_fast* net_b,
                                     const float* f_norm, const float* f_denorm,
                                     float* sigma, float* rgb, float* seg)
{
    float y[256];
    if (kind == NFO_DEC_OSG || kind == NFO_DEC_SEGMENTATION) {
        nfo_mlp_eval_fast(net_a, f_denorm, y);
        *sigma = y[0];
        for (int c = 1; c < net_a->out_dim; ++c) rgb[c - 1] = nfo_sigmoid(y[c]) * 1.002f - 0.001f;
        if (kind == NFO_DEC_SEGMENTATION) {
            nfo_mlp_eval_fast(net_b, f_denorm, y);
            for (int c = 0; c < net_b->out_dim; ++c) seg[c] = y[c];
        }
    } else {
        nfo_mlp_eval_fast(net_a, f_norm, y);     /* geo_net */
        *sigma = y[0];
        for (int c = 1; c < net_a->out_dim; ++c) seg[c - 1] = y[c];
        nfo_mlp_eval_fast(net_b, f_denorm, y);   /* app_net */
        for (int c = 0; c < net_b->out_dim; ++c) rgb[c] = nfo_sigmoid(y[c]) * 1.002f - 0.001f;
    }
}

static inline void nfo_plane_mean(const float* f0, const float* f1, const float* f2, int C, float* out)
{
    for (int c = 0; c < C; ++c) out[c] = ((f0[c] + f1[c]) + f2[c]) / 3.0f;
}

/* decoder(sampled_features[N,3,M,C], ...) stand-alone; feat_norm may be NULL for OSG. */
NFO_API void nfo_decoder(int kind, const nfo_mlp* net_a, const nfo_mlp* net_b,
                         const float* feat_norm, const float* feat_denorm,
                         int n, int64_t m, int C, int color_dim, int seg_dim,
                         float* rgb /*[N,M,color]*/, float* sigma /*[N,M]*/, float* seg /*[N,M,seg]*/)
{
    nfo_mlp_fast* fa = (nfo_mlp_fast*)malloc(2 * sizeof(nfo_mlp_fast));
    nfo_mlp_fast* fb = fa + 1;
    if (nfo_mlp_prepare(net_a, fa) || nfo_mlp_prepare(net_b, fb)) { free(fa); return; }
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < n; ++b) {
        for (int64_t i = 0; i < m; ++i) {
            float fn[64], fd[64];
            const int64_t base = (int64_t)b * 3 * m;
            if (feat_norm)
                nfo_plane_mean(feat_norm + (base + i) * C, feat_norm + (base + m + i) * C,
                               feat_norm + (base + 2 * m + i) * C, C, fn);
            nfo_plane_mean(feat_denorm + (base + i) * C, feat_denorm + (base + m + i) * C,
                           feat_denorm + (base + 2 * m + i) * C, C, fd);
            const int64_t s = (int64_t)b * m + i;
            nfo_decode_sample(kind, fa, fb, fn, fd, sigma + s, rgb + s * color_dim,
                              seg ? seg + s * seg_dim : NULL);
        }
    }
    free(fa);
}

/* ------------------------------------------------------------------------------------------
 * Ray marcher — training/volumetric_rendering/ray_marcher.py:25-57 (colour) and :68-101 (+seg)
 * One ray.  S samples -> S-1 mid-point intervals.  Transmittance is an exclusive cumprod of
 * (1 - alpha + 1e-10); torch-CPU cumprod accumulates fp32 data in double, so does this.
 * Returns unclamped depth (sum(w*t_mid)/sum(w), NaN when sum(w)==0); the caller applies
 * nan_to_num(+inf) and the GLOBAL [min(depths), max(depths)] clamp (ray_marcher.py:49-50,93-94).
 * ---------------------------------------------------------------------------------------- */
static void nfo_march_ray(const float* colors, int cc, const float* segs, int cs,
                          const float* sigma, const float* depth, int S, int white_back,
                          float* rgb, float* seg, float* depth_out, float* weights /*[S-1]*/, float* wsum)
{
    for (int c = 0; c < cc; ++c) rgb[c] = 0.0f;
    for (int c = 0; c < cs; ++c) seg[c] = 0.0f;
    double T = 1.0;
    float wd = 0.0f, wt = 0.0f;
    for (int i = 0; i < S - 1; ++i) {
        const float delta = depth[i + 1] - depth[i];
        const float sig_mid = (sigma[i] + sigma[i + 1]) / 2.0f;
        const float d_mid = (depth[i] + depth[i + 1]) / 2.0f;
        const float dens = nfo_softplus(sig_mid - 1.0f);
        const float alpha = 1.0f - expf(-(dens * delta));
        const float w = alpha * (float)T;
        T *= (double)((1.0f - alpha) + 1e-10f);
        weights[i] = w;
        for (int c = 0; c < cc; ++c) rgb[c] += w * ((colors[i * cc + c] + colors[(i + 1) * cc + c]) / 2.0f);
        for (int c = 0; c < cs; ++c) seg[c] += w * ((segs[i * cs + c] + segs[(i + 1) * cs + c]) / 2.0f);
        wd += w * d_mid;
        wt += w;
    }
    *depth_out = wd / wt;
    *wsum = wt;
    for (int c = 0; c < cc; ++c) {
        float v = rgb[c];
        if (white_back) v = v + 1.0f - wt;
        rgb[c] = v * 2.0f - 1.0f;
    }
}

static inline float nfo_finish_depth(float d, float dmin, float dmax)
{
    if (d != d) d = INFINITY;                 /* nan_to_num(x, nan=inf) ...             */
    if (d == INFINITY) d = FLT_MAX;           /* ... whose posinf/neginf default to ±max */
    if (d == -INFINITY) d = -FLT_MAX;
    if (d < dmin) d = dmin;
    if (d > dmax) d = dmax;
    return d;
}

NFO_API void nfo_ray_march(const float* colors /*[Rn,S,cc]*/, const float* segs /*[Rn,S,cs] or NULL*/,
                           const float* sigma /*[Rn,S]*/, const float* depths /*[Rn,S]*/,
                           int64_t n_rays, int S, int cc, int cs, int white_back,
                           float* rgb, float* seg, float* depth_out, float* weights /*[Rn,S-1]*/)
{
    float dmin = INFINITY, dmax = -INFINITY;
    for (int64_t i = 0; i < n_rays * S; ++i) {
        if (depths[i] < dmin) dmin = depths[i];
        if (depths[i] > dmax) dmax = depths[i];
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) {
        float segbuf[64], wsum, d;
        nfo_march_ray(colors + r * S * cc, cc, segs ? segs + r * S * cs : NULL, segs ? cs : 0,
                      sigma + r * S, depths + r * S, S, white_back,
                      rgb + r * cc, segs ? seg + r * cs : segbuf, &d, weights + r * (S - 1), &wsum);
        depth_out[r] = nfo_finish_depth(d, dmin, dmax);
    }
}

/* ------------------------------------------------------------------------------------------
 * Importance resampling — training/volumetric_rendering/renderer.py:194-253
 *   weights[S-1] -> max_pool1d(2,1,pad=1) [S] -> avg_pool1d(2,1) [S-1] -> +0.01
 *   bins = mid-depths [S-1];  sample_pdf(bins, smoothed[1:-1] (S-3 entries), S_f):
 *   pdf = (w+1e-5)/sum;  cdf = [0, cumsum(pdf)] (S-2 entries);  searchsorted(right=True);
 *   below = max(i-1,0), above = min(i, S-3);  denom<1e-5 -> 1.
 * NORMATIVE arithmetic: the normaliser is the exact (double) sum rounded to
 * fp32, the CDF a double running sum rounded per entry (torch-CPU cumsum does exactly that;
 * torch.sum's fp32 order is ISA-dependent, which is why the normaliser is pinned here).
 * `u` is [S_f] shared by all rays (u_per_ray == 0) or [Rn,S_f].
 * ---------------------------------------------------------------------------------------- */
static void nfo_resample_ray(const float* z, const float* w, int S, int s_f, const float* u,
                             float* out, int32_t* below_out, int32_t* above_out)
{
    float m[512], a[512], bins[512], cdf[512];
    const int nw = S - 1;
    m[0] = w[0];
    for (int i = 1; i < nw; ++i) m[i] = w[i - 1] > w[i] ? w[i - 1] : w[i];
    m[nw] = w[nw - 1];
    for (int i = 0; i < nw; ++i) a[i] = (m[i] + m[i + 1]) / 2.0f + 0.01f;
    for (int i = 0; i < nw; ++i) bins[i] = 0.5f * (z[i] + z[i + 1]);
    const int ns = S - 3;                       /* N_samples_ */
    double tot = 0.0;
    for (int j = 0; j < ns; ++j) tot += (double)(a[j + 1] + 1e-5f);
    const float totf = (float)tot;
    double run = 0.0;
    cdf[0] = 0.0f;
    for (int j = 0; j < ns; ++j) {
        const float pdf = (a[j + 1] + 1e-5f) / totf;
        run += (double)pdf;
        cdf[j + 1] = (float)run;
    }
    for (int k = 0; k < s_f; ++k) {
        const float uk = u[k];
        int ind = 0;                            /* first index with cdf[ind] > u  (right=True) */
        while (ind < ns + 1 && cdf[ind] <= uk) ++ind;
        const int below = ind - 1 < 0 ? 0 : ind - 1;
        const int above = ind > ns ? ns : ind;
        float den = cdf[above] - cdf[below];
        if (den < 1e-5f) den = 1.0f;
        const float frac = (uk - cdf[below]) / den;
        out[k] = bins[below] + frac * (bins[above] - bins[below]);
        if (below_out) below_out[k] = below;
        if (above_out) above_out[k] = above;
    }
}

NFO_API void nfo_importance_resample(const float* z_vals /*[Rn,S]*/, const float* weights /*[Rn,S-1]*/,
                                     int64_t n_rays, int S, int s_f, const float* u, int u_per_ray,
                                     float* out /*[Rn,S_f]*/, int32_t* below /*opt*/, int32_t* above /*opt*/)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r)
        nfo_resample_ray(z_vals + r * S, weights + r * (S - 1), S, s_f,
                         u_per_ray ? u + r * s_f : u, out + r * s_f,
                         below ? below + r * s_f : NULL, above ? above + r * s_f : NULL);
}

/* ------------------------------------------------------------------------------------------
 * Merge of coarse and fine samples — renderer.py:150-167,288-300
 *   cat(coarse, fine) then sort by depth.  order[k] = index into the concatenation.  The sort
 *   here is stable (ties keep coarse first); torch.sort is not flagged stable, but tied
 *   samples have identical positions and therefore identical attributes.
 * ---------------------------------------------------------------------------------------- */
static void nfo_merge_order(const float* d, int n, int32_t* order)
{
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 1; i < n; ++i) {               /* insertion sort: stable, n <= 768 */
        const int32_t k = order[i];
        int j = i - 1;
        while (j >= 0 && d[order[j]] > d[k]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = k;
    }
}

NFO_API void nfo_unify_order(const float* depths_cat /*[Rn,S]*/, int64_t n_rays, int S, int32_t* order)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) nfo_merge_order(depths_cat + r * S, S, order + r * S);
}

/* ------------------------------------------------------------------------------------------
 * Full forward — renderer.py:88-148 (ImportanceRenderer) and :301-363 (Disentangled...)
 * Planes arrive in the reference's NCHW layout [N,3,C,H,W] (plane_batch may be 1 to share one
 * plane set across the ray batch).  Per ray: coarse depths -> gather+decode -> coarse
 * weights -> resample -> gather+decode -> merge -> composite.  The depth clamp needs the
 * min/max of ALL merged depths, so unclamped depths are finished in a second sweep.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int kind;                 /* NFO_DEC_* */
    int C, H, W;              /* plane channels / size */
    int s_c, s_f;             /* depth_resolution, depth_resolution_importance (0 = single pass) */
    int color_dim, seg_dim;   /* 32, 15 (seg_dim 0 for OSG) */
    int white_back;
    float box_warp;
    float density_noise;      /* must be 0 here: the oracle has no RNG */
} nfo_render_cfg;

static void nfo_eval_point(const nfo_render_cfg* cfg, const nfo_mlp_fast* net_a, const nfo_mlp_fast* net_b,
                           const float* pl_norm /*channel-last [3,H,W,C] or NULL*/, const float* pl_denorm,
                           const float x[3], float* sigma, float* rgb, float* seg)
{
    const float scale = (float)(2.0 / (double)cfg->box_warp);
    const float q[3] = { scale * x[0], scale * x[1], scale * x[2] };
    float f[3][64], fn[64], fd[64];
    const int64_t plane_sz = (int64_t)cfg->H * cfg->W * cfg->C;
    for (int set = 0; set < 2; ++set) {
        const float* pl = set == 0 ? pl_norm : pl_denorm;
        if (!pl) continue;
        for (int p = 0; p < 3; ++p) {
            float gx, gy; nfo_project(q, p, &gx, &gy);
            nfo_gather_plane(pl + p * plane_sz, cfg->C, cfg->H, cfg->W, 1, (int64_t)cfg->W * cfg->C, cfg->C,
                             gx, gy, f[p]);
        }
        nfo_plane_mean(f[0], f[1], f[2], cfg->C, set == 0 ? fn : fd);
    }
    nfo_decode_sample(cfg->kind, net_a, net_b, fn, fd, sigma, rgb, seg);
}

static float* nfo_to_channel_last(const float* planes, int64_t n_img, int C, int64_t hw)
{
    float* out = (float*)malloc(sizeof(float) * (size_t)(n_img * C * hw));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_img; ++i)
        for (int64_t px = 0; px < hw; ++px)
            for (int c = 0; c < C; ++c)
                out[(i * hw + px) * C + c] = planes[(i * C + c) * hw + px];
    return out;
}

NFO_API int nfo_render(const nfo_render_cfg* cfg, const nfo_mlp* net_a, const nfo_mlp* net_b,
                       const float* planes_norm /*[Np,3,C,H,W] or NULL (OSG)*/,
                       const float* planes_denorm /*[Np,3,C,H,W]*/, int plane_batch,
                       const float* origins /*[N,R,3]*/, const float* dirs /*[N,R,3]*/, int n, int64_t n_rays,
                       const float* depths_coarse /*[N,R,S_c] (already jittered)*/,
                       const float* u_fine, int u_per_ray,
                       float* rgb /*[N,R,color]*/, float* seg /*[N,R,seg] or NULL*/,
                       float* depth /*[N,R]*/, float* wsum /*[N,R]*/,
                       float* depths_fine_out /*[N,R,S_f] or NULL*/, float* weights_coarse_out /*[N,R,S_c-1] or NULL*/)
{
    const int S_c = cfg->s_c, S_f = cfg->s_f, S_all = S_c + S_f;
    const int cc = cfg->color_dim, cs = cfg->seg_dim;
    if (S_all > 768 || cc > 64 || cs > 64 || cfg->C > 64) return 1;
    nfo_mlp_fast* fa = (nfo_mlp_fast*)malloc(2 * sizeof(nfo_mlp_fast));
    nfo_mlp_fast* fb = fa + 1;
    if (nfo_mlp_prepare(net_a, fa) || nfo_mlp_prepare(net_b, fb)) { free(fa); return 1; }
    const int64_t hw = (int64_t)cfg->H * cfg->W;
    float* cl_norm = planes_norm ? nfo_to_channel_last(planes_norm, (int64_t)plane_batch * 3, cfg->C, hw) : NULL;
    float* cl_denorm = nfo_to_channel_last(planes_denorm, (int64_t)plane_batch * 3, cfg->C, hw);
    const int64_t set_sz = 3 * hw * cfg->C;
    const int64_t total = (int64_t)n * n_rays;
    float gmin = INFINITY, gmax = -INFINITY;

#pragma omp parallel for schedule(dynamic, 16) reduction(min : gmin) reduction(max : gmax)
    for (int64_t ray = 0; ray < total; ++ray) {
        const int b = (int)(ray / n_rays);
        const int pb = plane_batch == 1 ? 0 : b;
        const float* pn = cl_norm ? cl_norm + pb * set_sz : NULL;
        const float* pd = cl_denorm + pb * set_sz;
        const float* o = origins + ray * 3;
        const float* d = dirs + ray * 3;
        float* t_all = (float*)malloc(sizeof(float) * (size_t)S_all * (2 + cc + cs + 2) + sizeof(int32_t) * S_all);
        float* sg_all = t_all + S_all;
        float* col_all = sg_all + S_all;
        float* seg_all = col_all + (size_t)S_all * cc;
        float* t_sorted = seg_all + (size_t)S_all * cs;    /* reused below */
        float wcoarse[768];
        for (int s = 0; s < S_c; ++s) {
            const float t = depths_coarse[ray * S_c + s];
            t_all[s] = t;
            const float x[3] = { o[0] + t * d[0], o[1] + t * d[1], o[2] + t * d[2] };
            nfo_eval_point(cfg, fa, fb, pn, pd, x, sg_all + s, col_all + (size_t)s * cc, seg_all + (size_t)s * cs);
        }
        int S = S_c;
        float rgbv[64], segv[64], dep, ws;
        if (S_f > 0) {
            nfo_march_ray(col_all, cc, seg_all, cs, sg_all, t_all, S_c, cfg->white_back, rgbv, segv, &dep, wcoarse, &ws);
            if (weights_coarse_out) memcpy(weights_coarse_out + ray * (S_c - 1), wcoarse, sizeof(float) * (S_c - 1));
            nfo_resample_ray(t_all, wcoarse, S_c, S_f, u_per_ray ? u_fine + ray * S_f : u_fine, t_all + S_c, NULL, NULL);
            if (depths_fine_out) memcpy(depths_fine_out + ray * S_f, t_all + S_c, sizeof(float) * S_f);
            for (int s = S_c; s < S_all; ++s) {
                const float t = t_all[s];
                const float x[3] = { o[0] + t * d[0], o[1] + t * d[1], o[2] + t * d[2] };
                nfo_eval_point(cfg, fa, fb, pn, pd, x, sg_all + s, col_all + (size_t)s * cc, seg_all + (size_t)s * cs);
            }
            S = S_all;
            /* unify_samples: permute everything into depth order */
            int32_t* order = (int32_t*)(t_sorted + S_all * 2);
            nfo_merge_order(t_all, S, order);
            float* tmp = (float*)malloc(sizeof(float) * (size_t)S * (2 + cc + cs));
            float* ts = tmp; float* ss = ts + S; float* csrt = ss + S; float* gsrt = csrt + (size_t)S * cc;
            for (int k = 0; k < S; ++k) {
                const int src = order[k];
                ts[k] = t_all[src]; ss[k] = sg_all[src];
                memcpy(csrt + (size_t)k * cc, col_all + (size_t)src * cc, sizeof(float) * cc);
                if (cs) memcpy(gsrt + (size_t)k * cs, seg_all + (size_t)src * cs, sizeof(float) * cs);
            }
            memcpy(t_all, ts, sizeof(float) * S); memcpy(sg_all, ss, sizeof(float) * S);
            memcpy(col_all, csrt, sizeof(float) * (size_t)S * cc);
            if (cs) memcpy(seg_all, gsrt, sizeof(float) * (size_t)S * cs);
            free(tmp);
        }
        float wfinal[768];
        nfo_march_ray(col_all, cc, seg_all, cs, sg_all, t_all, S, cfg->white_back, rgbv, segv, &dep, wfinal, &ws);
        memcpy(rgb + ray * cc, rgbv, sizeof(float) * cc);
        if (seg && cs) memcpy(seg + ray * cs, segv, sizeof(float) * cs);
        depth[ray] = dep;      /* unclamped; finished below */
        wsum[ray] = ws;
        for (int k = 0; k < S; ++k) { if (t_all[k] < gmin) gmin = t_all[k]; if (t_all[k] > gmax) gmax = t_all[k]; }
        free(t_all);
    }
    for (int64_t ray = 0; ray < total; ++ray) depth[ray] = nfo_finish_depth(depth[ray], gmin, gmax);
    free(cl_norm); free(cl_denorm); free(fa);
    return 0;
}

/* run_model — renderer.py:142-148,259-287: gather + decode at explicit points, no marching. */
NFO_API int nfo_run_model(const nfo_render_cfg* cfg, const nfo_mlp* net_a, const nfo_mlp* net_b,
                          const float* planes_norm, const float* planes_denorm, int plane_batch,
                          const float* coords /*[N,M,3]*/, int n, int64_t m,
                          float* rgb, float* sigma, float* seg)
{
    if (cfg->color_dim > 64 || cfg->seg_dim > 64 || cfg->C > 64) return 1;
    nfo_mlp_fast* fa = (nfo_mlp_fast*)malloc(2 * sizeof(nfo_mlp_fast));
    nfo_mlp_fast* fb = fa + 1;
    if (nfo_mlp_prepare(net_a, fa) || nfo_mlp_prepare(net_b, fb)) { free(fa); return 1; }
    const int64_t hw = (int64_t)cfg->H * cfg->W;
    float* cl_norm = planes_norm ? nfo_to_channel_last(planes_norm, (int64_t)plane_batch * 3, cfg->C, hw) : NULL;
    float* cl_denorm = nfo_to_channel_last(planes_denorm, (int64_t)plane_batch * 3, cfg->C, hw);
    const int64_t set_sz = 3 * hw * cfg->C;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n * m; ++i) {
        const int pb = plane_batch == 1 ? 0 : (int)(i / m);
        float segv[64];
        nfo_eval_point(cfg, fa, fb, cl_norm ? cl_norm + pb * set_sz : NULL, cl_denorm + pb * set_sz,
                       coords + i * 3, sigma + i, rgb + i * cfg->color_dim, seg ? seg + i * cfg->seg_dim : segv);
    }
    free(cl_norm); free(cl_denorm); free(fa);
    return 0;
}

NFO_API int nfo_version(void) { return 1; }


/* ------------------------------------------------------------------------------------------
 * SR pre-resize: torch.nn.functional.interpolate(x, size, mode='bilinear',
 * align_corners=False, antialias=sr_antialias) as called on the rendered feature image and its first three
 * channels at training/superresolution.py:48-52,80-84,282-286.  The arithmetic lives in ATen (not under
 * /root/reference); restated from its published definition:
 *   scale = in/out per axis;
 *   antialias: triangle filter of half-width support = max(scale, 1) around center = scale*(o + 0.5), taps
 *     [max(int(center - support + 0.5), 0), min(int(center + support + 0.5), in)), weight
 *     max(0, 1 - |(j - center + 0.5) / max(scale, 1)|), normalised to sum 1; horizontal pass, then vertical;
 *   plain: src = max(scale*(o + 0.5) - 0.5, 0), i0 = floor(src), i1 = min(i0 + 1, in - 1), weights (1 - t, t).
 * in [n_img, ih, iw] -> out [n_img, oh, ow]. */
static int nfo_aa_weights(int o, int in, int out, int antialias, int* first, float* w /* >= 2*ceil(max(scale,1)) + 2 */)
{
    const float scale = (float)in / (float)out;
    if (!antialias) {
        float src = scale * ((float)o + 0.5f) - 0.5f;
        if (src < 0.0f) src = 0.0f;
        int i0 = (int)src;
        if (i0 > in - 1) i0 = in - 1;
        const int has1 = i0 < in - 1;
        const float t = src - (float)i0;
        *first = i0;
        w[0] = 1.0f - t;
        w[1] = t;
        if (!has1) { w[0] = 1.0f; return 1; }      /* i1 == i0: both weights land on the same texel */
        return 2;
    }
    const float support = scale >= 1.0f ? scale : 1.0f;
    const float inv = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const float center = scale * ((float)o + 0.5f);
    int lo = (int)(center - support + 0.5f);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5f);
    if (hi > in) hi = in;
    const int n = hi - lo;
    float total = 0.0f;
    for (int j = 0; j < n; ++j) {
        float x = ((float)(j + lo) - center + 0.5f) * inv;
        if (x < 0.0f) x = -x;
        w[j] = x < 1.0f ? 1.0f - x : 0.0f;
        total += w[j];
    }
    for (int j = 0; j < n; ++j) w[j] = total != 0.0f ? w[j] / total : 0.0f;
    *first = lo;
    return n;
}

NFO_API void nfo_resize_bilinear(const float* in, int64_t n_img, int ih, int iw, int oh, int ow, int antialias, float* out)
{
    const int max_taps_x = 2 * (int)ceilf((float)iw / (float)ow > 1.0f ? (float)iw / (float)ow : 1.0f) + 3;
    const int max_taps_y = 2 * (int)ceilf((float)ih / (float)oh > 1.0f ? (float)ih / (float)oh : 1.0f) + 3;
#pragma omp parallel
    {
        float* wx = (float*)malloc(sizeof(float) * (size_t)max_taps_x);
        float* wy = (float*)malloc(sizeof(float) * (size_t)max_taps_y);
        float* row = (float*)malloc(sizeof(float) * (size_t)ih * (size_t)ow);     /* horizontal pass of one image */
#pragma omp for schedule(static)
        for (int64_t img = 0; img < n_img; ++img) {
            const float* src = in + img * ih * iw;
            for (int x = 0; x < ow; ++x) {
                int x0; const int nx = nfo_aa_weights(x, iw, ow, antialias, &x0, wx);
                for (int y = 0; y < ih; ++y) {
                    float acc = 0.0f;
                    for (int j = 0; j < nx; ++j) acc += wx[j] * src[y * iw + x0 + j];
                    row[y * ow + x] = acc;
                }
            }
            for (int y = 0; y < oh; ++y) {
                int y0; const int ny = nfo_aa_weights(y, ih, oh, antialias, &y0, wy);
                for (int x = 0; x < ow; ++x) {
                    float acc = 0.0f;
                    for (int j = 0; j < ny; ++j) acc += wy[j] * row[(y0 + j) * ow + x];
                    out[img * oh * ow + y * ow + x] = acc;
                }
            }
        }
        free(wx); free(wy); free(row);
    }
}


/* Thread control for the timing legs of bench.py: launchers such as torchrun export OMP_NUM_THREADS=1, which libgomp
 * reads once at load time; this sets the team size explicitly.  Returns the previous maximum. */
#ifdef _OPENMP
#include <omp.h>
#endif
NFO_API int nfo_set_num_threads(int n)
{
#ifdef _OPENMP
    const int prev = omp_get_max_threads();
    if (n > 0) omp_set_num_threads(n);
    return prev;
#else
    (void)n;
    return 1;
#endif
}
