// Backbone / super-resolution elementwise plugins (first step): the two custom ops every StyleGAN2 layer of the
// reference calls around its convolution,
//
//   bias_act    torch_utils/ops/bias_act.py:51-209 (plugin bias_act.cpp:34-99, bias_act.cu:27-151): y = clamp(gain * act(x + b[dim])),
//               its first derivative (grad = 1: dx from dy and the saved x / y) and second derivative (grad = 2), 9 activations
//   upfirdn2d   torch_utils/ops/upfirdn2d.py:117-214 (plugin upfirdn2d.cpp:20-107, upfirdn2d.cu:33-204): zero-insert upsampling,
//               zero padding / cropping, FIR filtering, decimation, in one pass
//
// Both are memory-bound: bias_act moves 16 bytes per thread and access with the bias index computed once per vector; the FIR
// fast path (up = down = 1, rows contiguous: what conv2d_resample runs after every transposed convolution, conv2d_resample.py:128-129)
// stages the input tile in shared memory and gives every thread a 4 x 2 block of outputs, so an output costs ~1.3 shared-memory
// 16-byte reads instead of fw*fh L1 reads; every other configuration (the 3-channel image up-sampling of the skip connection,
// strided layouts) takes the generic one-thread-per-output kernel, which only visits the taps that land on a real input sample.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "nfe_common.cuh"

namespace nfe {

namespace sg {

template <class T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------- bias_act
constexpr float SELU_SCALE = 1.0507009873554804934193349852946f;
constexpr float SELU_ALPHA = 1.6732632423543772848170429916717f;
constexpr float EXP_RANGE = 80.0f;

// One element.  G == 0: x is the input, returns clamp(gain * act(x + b)).  G == 1: x is dy, returns d/dx of the forward at the saved
// point times dy; G == 2: x is the incoming second-order gradient, dy the first-order one.  xref = saved input + b (only swish needs
// it), yref = saved forward output.  Same case analysis as the reference kernel (bias_act.cu:56-135), written from the formulas.
template <int A>
__device__ __forceinline__ float bias_act_one(int G, float x, float b, float xref, float yref, float dy, float alpha, float gain, float clamp)
{
    if (G == 0) x += b; else xref += b;
    const float yy = gain != 0.0f ? yref / gain : 0.0f;
    float y = 0.0f;
    if (A == 1) { if (G <= 1) y = x; }
    if (A == 2) { if (G == 0) y = x > 0.0f ? x : 0.0f; if (G == 1) y = yy > 0.0f ? x : 0.0f; }
    if (A == 3) { if (G == 0) y = x > 0.0f ? x : x * alpha; if (G == 1) y = yy > 0.0f ? x : x * alpha; }
    if (A == 4) {
        if (G == 0) y = tanhf(x);
        if (G == 1) y = x * (1.0f - yy * yy);
        if (G == 2) y = x * (1.0f - yy * yy) * (-2.0f * yy);
    }
    if (A == 5) {
        if (G == 0) y = x < -EXP_RANGE ? 0.0f : 1.0f / (expf(-x) + 1.0f);
        if (G == 1) y = x * yy * (1.0f - yy);
        if (G == 2) y = x * yy * (1.0f - yy) * (1.0f - 2.0f * yy);
    }
    if (A == 6) {
        if (G == 0) y = x >= 0.0f ? x : expm1f(x);
        if (G == 1) y = yy >= 0.0f ? x : x * (yy + 1.0f);
        if (G == 2) y = yy >= 0.0f ? 0.0f : x * (yy + 1.0f);
    }
    if (A == 7) {
        if (G == 0) y = x >= 0.0f ? SELU_SCALE * x : (SELU_SCALE * SELU_ALPHA) * expm1f(x);
        if (G == 1) y = yy >= 0.0f ? x * SELU_SCALE : x * (yy + SELU_SCALE * SELU_ALPHA);
        if (G == 2) y = yy >= 0.0f ? 0.0f : x * (yy + SELU_SCALE * SELU_ALPHA);
    }
    if (A == 8) {
        if (G == 0) y = x > 20.0f ? x : log1pf(expf(x));            // torch.nn.functional.softplus (threshold 20), bias_act.py:32
        if (G == 1) y = x * (1.0f - expf(-yy));
        if (G == 2) { const float c = expf(-yy); y = x * c * (1.0f - c); }
    }
    if (A == 9) {
        if (G == 0) y = x < -EXP_RANGE ? 0.0f : x / (expf(-x) + 1.0f);
        else {
            const float c = expf(xref), d = c + 1.0f;
            if (G == 1) y = xref > 0.5f * EXP_RANGE ? x : x * c * (xref + d) / (d * d);
            else y = xref > 0.5f * EXP_RANGE ? 0.0f : x * c * (xref * (2.0f - d) + 2.0f * d) / (d * d * d);
            yref = xref < -EXP_RANGE ? 0.0f : xref / (expf(-xref) + 1.0f) * gain;
        }
    }
    y *= gain * dy;
    if (clamp >= 0.0f) {
        if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0.0f ? clamp : -clamp);
        else y = (yref > -clamp && yref < clamp) ? y : 0.0f;
    }
    return y;
}

struct BiasActArgs {
    const void *x, *b, *xref, *yref, *dy;
    void* y;
    int64_t size_x;
    int size_b;
    int64_t step_b;
    int grad;
    float alpha, gain, clamp;
};

template <class T> struct Vec { static constexpr int N = 16 / sizeof(T); };

// 16 bytes per access.  MODE 0: all elements of a vector share one bias entry (step_b % N == 0: NCHW with H*W % N == 0);
// MODE 1: consecutive elements take consecutive bias entries (step_b == 1 and size_b % N == 0: channels-last / [M, C] activations);
// MODE 2: scalar fallback (any shape, any alignment).
template <class T, int A, int MODE>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActArgs p)
{
    constexpr int N = Vec<T>::N;
    const T* x = static_cast<const T*>(p.x);
    const T* b = static_cast<const T*>(p.b);
    const T* xref = static_cast<const T*>(p.xref);
    const T* yref = static_cast<const T*>(p.yref);
    const T* dy = static_cast<const T*>(p.dy);
    T* y = static_cast<T*>(p.y);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (MODE == 2) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.size_x; i += stride) {
            const float bv = b ? to_f(b[(i / p.step_b) % p.size_b]) : 0.0f;
            y[i] = from_f<T>(bias_act_one<A>(p.grad, to_f(x[i]), bv, xref ? to_f(xref[i]) : 0.0f, yref ? to_f(yref[i]) : 0.0f,
                                             dy ? to_f(dy[i]) : 1.0f, p.alpha, p.gain, p.clamp));
        }
        return;
    }
    struct alignas(16) Pack { T v[N]; };
    const int64_t n_vec = p.size_x / N;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += stride) {
        const int64_t i0 = v * N;
        const Pack px = reinterpret_cast<const Pack*>(x)[v];
        Pack pr, pyr, pdy, pb, out;
        if (xref) pr = reinterpret_cast<const Pack*>(xref)[v];
        if (yref) pyr = reinterpret_cast<const Pack*>(yref)[v];
        if (dy) pdy = reinterpret_cast<const Pack*>(dy)[v];
        float b0 = 0.0f;
        if (b) {
            if (MODE == 0) b0 = to_f(b[(i0 / p.step_b) % p.size_b]);
            else pb = *reinterpret_cast<const Pack*>(b + i0 % p.size_b);
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float bv = (b && MODE == 1) ? to_f(pb.v[k]) : b0;
            out.v[k] = from_f<T>(bias_act_one<A>(p.grad, to_f(px.v[k]), bv, xref ? to_f(pr.v[k]) : 0.0f, yref ? to_f(pyr.v[k]) : 0.0f,
                                                 dy ? to_f(pdy.v[k]) : 1.0f, p.alpha, p.gain, p.clamp));
        }
        reinterpret_cast<Pack*>(y)[v] = out;
    }
}

template <class T, int A>
int bias_act_launch(const BiasActArgs& p, cudaStream_t stream)
{
    constexpr int N = Vec<T>::N;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec_ok = p.size_x % N == 0 && aligned(p.x) && aligned(p.xref) && aligned(p.yref) && aligned(p.dy) && aligned(p.y);
    int mode = 2;
    if (vec_ok && (!p.b || p.step_b % N == 0)) mode = 0;
    else if (vec_ok && p.step_b == 1 && p.size_b % N == 0 && aligned(p.b)) mode = 1;
    const int64_t work = mode == 2 ? p.size_x : p.size_x / N;
    const int blocks = (int)std::min<int64_t>((work + 255) / 256, (int64_t)sm_count() * 8);
    if (mode == 0) bias_act_kernel<T, A, 0><<<blocks, 256, 0, stream>>>(p);
    else if (mode == 1) bias_act_kernel<T, A, 1><<<blocks, 256, 0, stream>>>(p);
    else bias_act_kernel<T, A, 2><<<blocks, 256, 0, stream>>>(p);
    return check_launch("bias_act_kernel");
}

template <class T>
int bias_act_dispatch(int act, const BiasActArgs& p, cudaStream_t stream)
{
    switch (act) {
    case 1: return bias_act_launch<T, 1>(p, stream);
    case 2: return bias_act_launch<T, 2>(p, stream);
    case 3: return bias_act_launch<T, 3>(p, stream);
    case 4: return bias_act_launch<T, 4>(p, stream);
    case 5: return bias_act_launch<T, 5>(p, stream);
    case 6: return bias_act_launch<T, 6>(p, stream);
    case 7: return bias_act_launch<T, 7>(p, stream);
    case 8: return bias_act_launch<T, 8>(p, stream);
    case 9: return bias_act_launch<T, 9>(p, stream);
    }
    set_error("nfe_bias_act: act must be 1..9 (bias_act.py:23-33 cuda_idx), got %d", act);
    return 1;
}

// ---------------------------------------------------------------------------------------------- upfirdn2d
struct UpfirdnArgs {
    const void* x;
    const float* f;
    void* y;
    int n, c, in_h, in_w, out_h, out_w, fh, fw;
    int64_t xs[4], ys[4];                 // element strides of x and y: batch, channel, row, column
    int upx, upy, downx, downy, padx0, pady0;
    int flip;
    float gain;
};

// y[oy,ox] = gain * sum_{ky,kx} F[ky,kx] * u[oy*downy + ky, ox*downx + kx], u = zero-inserted, zero-padded x, F = f flipped unless
// flip_filter (upfirdn2d.py:169-214).  Only the taps that land on a real sample are visited.
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(UpfirdnArgs p)
{
    const int64_t total = (int64_t)p.n * p.c * p.out_h * p.out_w;
    const T* x = static_cast<const T*>(p.x);
    T* y = static_cast<T*>(p.y);
    // the fastest-varying index of the thread order follows y's unit stride (NCHW: column, channels-last: channel)
    const bool cl = p.ys[1] == 1 && p.c > 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int ox, oy, ch, nb;
        int64_t r = i;
        if (cl) { ch = (int)(r % p.c); r /= p.c; ox = (int)(r % p.out_w); r /= p.out_w; oy = (int)(r % p.out_h); nb = (int)(r / p.out_h); }
        else { ox = (int)(r % p.out_w); r /= p.out_w; oy = (int)(r % p.out_h); r /= p.out_h; ch = (int)(r % p.c); nb = (int)(r / p.c); }
        const int my = oy * p.downy - p.pady0, mx = ox * p.downx - p.padx0;      // u row my + pady0 + ky holds x row (my + ky) / upy
        // first tap whose row / column is a multiple of the up-sampling factor
        int ky0 = ((-my) % p.upy + p.upy) % p.upy, kx0 = ((-mx) % p.upx + p.upx) % p.upx;
        const T* xb = x + nb * p.xs[0] + ch * p.xs[1];
        float acc = 0.0f;
        for (int ky = ky0; ky < p.fh; ky += p.upy) {
            const int iy = (my + ky) / p.upy;
            if (iy < 0 || iy >= p.in_h) continue;
            const int fy = p.flip ? ky : p.fh - 1 - ky;
            for (int kx = kx0; kx < p.fw; kx += p.upx) {
                const int ix = (mx + kx) / p.upx;
                if (ix < 0 || ix >= p.in_w) continue;
                const int fx = p.flip ? kx : p.fw - 1 - kx;
                acc = fmaf(to_f(xb[iy * p.xs[2] + ix * p.xs[3]]), __ldg(p.f + fy * p.fw + fx), acc);
            }
        }
        y[nb * p.ys[0] + ch * p.ys[1] + oy * p.ys[2] + ox * p.ys[3]] = from_f<T>(acc * p.gain);
    }
}

// FIR fast path: up = down = 1, unit column stride on both sides, filter at most 8 x 8.  A CTA owns a 16 x 128 output tile of one
// (batch, channel) plane; thread (tx, ty) computes the 2 rows x 4 columns at (2 ty, 4 tx) from the staged (16 + fh - 1) x (128 + 8)
// input tile with 16-byte shared-memory reads (three per input row cover 4 + 7 columns).
constexpr int FIR_TW = 128, FIR_TH = 16, FIR_MAXF = 8, FIR_PITCH = FIR_TW + 12;

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_fir_kernel(UpfirdnArgs p, int tiles_x, int tiles_y)
{
    __shared__ __align__(16) float tile[(FIR_TH + FIR_MAXF - 1) * FIR_PITCH];
    __shared__ float filt[FIR_MAXF * FIR_MAXF];
    const T* x = static_cast<const T*>(p.x);
    T* y = static_cast<T*>(p.y);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < p.fh * p.fw) {
        const int ky = threadIdx.x / p.fw, kx = threadIdx.x % p.fw;
        filt[threadIdx.x] = __ldg(p.f + (p.flip ? ky : p.fh - 1 - ky) * p.fw + (p.flip ? kx : p.fw - 1 - kx)) * p.gain;
    }
    const int64_t n_tiles = (int64_t)p.n * p.c * tiles_y * tiles_x;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int bx = (int)(t % tiles_x), by = (int)((t / tiles_x) % tiles_y);
        const int64_t plane = t / ((int64_t)tiles_x * tiles_y);
        const int ch = (int)(plane % p.c), nb = (int)(plane / p.c);
        const int ox0 = bx * FIR_TW, oy0 = by * FIR_TH;
        const T* xb = x + nb * p.xs[0] + ch * p.xs[1];
        const int rows = FIR_TH + p.fh - 1, cols = FIR_TW + p.fw - 1;
        __syncthreads();                                   // previous tile fully consumed (and the filter visible)
        for (int r = ty; r < rows; r += 8) {
            const int iy = oy0 + r - p.pady0;
            const bool row_ok = iy >= 0 && iy < p.in_h;
            for (int cc = tx; cc < cols; cc += 32) {
                const int ix = ox0 + cc - p.padx0;
                tile[r * FIR_PITCH + cc] = (row_ok && ix >= 0 && ix < p.in_w) ? to_f(xb[iy * p.xs[2] + ix]) : 0.0f;
            }
        }
        __syncthreads();
        float acc[2][4] = {};
        for (int r = 0; r < p.fh + 1; ++r) {                // input rows 2 ty + r, r = 0 .. fh: output row j uses filter row r - j
            const float4* src = reinterpret_cast<const float4*>(&tile[(2 * ty + r) * FIR_PITCH + 4 * tx]);
            const float4 a = src[0], b = src[1], c = src[2];
            const float in[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int ky = r - j;
                if (ky < 0 || ky >= p.fh) continue;
#pragma unroll
                for (int kx = 0; kx < FIR_MAXF; ++kx) {
                    if (kx >= p.fw) break;
                    const float w = filt[ky * p.fw + kx];
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(in[i + kx], w, acc[j][i]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int oy = oy0 + 2 * ty + j;
            if (oy >= p.out_h) continue;
            T* dst = y + nb * p.ys[0] + ch * p.ys[1] + oy * p.ys[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ox = ox0 + 4 * tx + i;
                if (ox < p.out_w) dst[ox] = from_f<T>(acc[j][i]);
            }
        }
    }
}

// Up-sampling by 2 with a filter of at most 4 x 4 (upsample2d of the skip image, networks_stylegan2.py:451; upfirdn2d.py:315-350),
// unit column stride on both sides: a thread owns a 2 x 2 block of outputs — every output then takes at most 2 x 2 taps, all index
// arithmetic is 32-bit and done once per block, and a warp's stores are contiguous.  (The generic kernel spends ~100 instructions per
// output on 64-bit index decoding and tap loops: 1.4 ms of a 7.5 ms backbone pass went there.)
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_up2_kernel(UpfirdnArgs p)
{
    __shared__ float filt[16];
    if ((int)threadIdx.x < p.fh * p.fw) {
        const int ky = threadIdx.x / p.fw, kx = threadIdx.x % p.fw;
        filt[threadIdx.x] = __ldg(p.f + (p.flip ? ky : p.fh - 1 - ky) * p.fw + (p.flip ? kx : p.fw - 1 - kx)) * p.gain;
    }
    __syncthreads();
    const int plane = blockIdx.z, nb = plane / p.c, ch = plane % p.c;
    const int bx = blockIdx.x * 64 + (threadIdx.x & 63), by = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int ox0 = 2 * bx, oy0 = 2 * by;
    if (ox0 >= p.out_w || oy0 >= p.out_h) return;
    const T* xb = static_cast<const T*>(p.x) + nb * p.xs[0] + ch * p.xs[1];
    T* yb = static_cast<T*>(p.y) + nb * p.ys[0] + ch * p.ys[1];
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int my = oy0 + dy - p.pady0, ky0 = my & 1;           // first tap whose (my + ky) is even: rows of real samples
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int ky = ky0 + 2 * a, iy = (my + ky) >> 1;       // (my + ky) is even: exact, also for negative values
            if (ky >= p.fh || iy < 0 || iy >= p.in_h) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int mx = ox0 + dx - p.padx0, kx0 = mx & 1;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int kx = kx0 + 2 * b, ix = (mx + kx) >> 1;
                    if (kx >= p.fw || ix < 0 || ix >= p.in_w) continue;
                    acc[dy][dx] = fmaf(to_f(xb[(int64_t)iy * p.xs[2] + ix]), filt[ky * p.fw + kx], acc[dy][dx]);
                }
            }
        }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        if (oy0 + dy >= p.out_h) continue;
        T* dst = yb + (int64_t)(oy0 + dy) * p.ys[2] + ox0;
        dst[0] = from_f<T>(acc[dy][0]);
        if (ox0 + 1 < p.out_w) dst[1] = from_f<T>(acc[dy][1]);
    }
}

template <class T>
int upfirdn2d_launch(const UpfirdnArgs& p, cudaStream_t stream)
{
    const int64_t total = (int64_t)p.n * p.c * p.out_h * p.out_w;
    if (total == 0) return 0;
    const bool fir = p.upx == 1 && p.upy == 1 && p.downx == 1 && p.downy == 1 && p.xs[3] == 1 && p.ys[3] == 1 && p.fw <= FIR_MAXF &&
                     p.fh <= FIR_MAXF && p.out_w >= 32;
    if (fir) {
        const int tiles_x = (p.out_w + FIR_TW - 1) / FIR_TW, tiles_y = (p.out_h + FIR_TH - 1) / FIR_TH;
        const int64_t n_tiles = (int64_t)p.n * p.c * tiles_x * tiles_y;
        const int blocks = (int)std::min<int64_t>(n_tiles, (int64_t)sm_count() * 8);
        upfirdn2d_fir_kernel<T><<<blocks, 256, 0, stream>>>(p, tiles_x, tiles_y);
        return check_launch("upfirdn2d_fir_kernel");
    }
    const bool up2 = p.upx == 2 && p.upy == 2 && p.downx == 1 && p.downy == 1 && p.xs[3] == 1 && p.ys[3] == 1 && p.fw <= 4 && p.fh <= 4 &&
                     (int64_t)p.n * p.c <= 65535;
    if (up2) {
        const dim3 grid((unsigned)((p.out_w + 127) / 128), (unsigned)((p.out_h + 7) / 8), (unsigned)(p.n * p.c));
        upfirdn2d_up2_kernel<T><<<grid, 256, 0, stream>>>(p);
        return check_launch("upfirdn2d_up2_kernel");
    }
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    upfirdn2d_generic_kernel<T><<<blocks, 256, 0, stream>>>(p);
    return check_launch("upfirdn2d_generic_kernel");
}


// ---------------------------------------------------------------------------------------------- layout conversion
// [n, c, hw] (contiguous NCHW) <-> [n, hw, c] (channels-last), same element type, through a 64 x 64 shared-memory tile so that both
// sides move whole 128-byte lines.  torch's own `.contiguous(memory_format=...)` runs these copies at a sixth of the memory
// bandwidth (0.54 ms for a 268 MB fp16 tensor, measured beside the convolution that consumes it in 0.53 ms).
template <class T>
__global__ void __launch_bounds__(256) layout_transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int rows, int cols, int tiles_r, int tiles_c)
{
    // src [n][rows][cols] -> dst [n][cols][rows]
    __shared__ T tile[64][64 + 4 / (int)sizeof(T)];          // odd pitch in 32-bit words: the transposed reads are conflict-free
    const int64_t n = blockIdx.y;
    const int tr = (blockIdx.x / tiles_c) * 64, tc0 = (blockIdx.x % tiles_c) * 64;
    const T* s = src + n * (int64_t)rows * cols;
    T* d = dst + n * (int64_t)rows * cols;
    const int lx = threadIdx.x & 63, ly = threadIdx.x >> 6;                 // 64 columns x 4 rows per pass
#pragma unroll 4
    for (int r = ly; r < 64; r += 4) {
        const int gr = tr + r, gc = tc0 + lx;
        if (gr < rows && gc < cols) tile[r][lx] = s[(int64_t)gr * cols + gc];
    }
    __syncthreads();
#pragma unroll 4
    for (int c = ly; c < 64; c += 4) {
        const int gc = tc0 + c, gr = tr + lx;
        if (gc < cols && gr < rows) d[(int64_t)gc * rows + gr] = tile[lx][c];
    }
}

// 2-byte elements, even rows / cols and 4-byte aligned bases: two elements per thread on both sides, so that a warp still moves 128 bytes
__global__ void __launch_bounds__(256) layout_transpose16_kernel(const unsigned short* __restrict__ src, unsigned short* __restrict__ dst, int rows, int cols,
                                                                  int tiles_r, int tiles_c)
{
    __shared__ unsigned short tile[64][64 + 2];
    const int64_t n = blockIdx.y;
    const int tr = (blockIdx.x / tiles_c) * 64, tc0 = (blockIdx.x % tiles_c) * 64;
    const unsigned short* s = src + n * (int64_t)rows * cols;
    unsigned short* d = dst + n * (int64_t)rows * cols;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;                 // 32 pairs x 8 rows per pass
#pragma unroll 4
    for (int r = ly; r < 64; r += 8) {
        const int gr = tr + r, gc = tc0 + 2 * lx;
        if (gr < rows && gc < cols) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(s + (int64_t)gr * cols + gc);
            tile[r][2 * lx] = (unsigned short)(v & 0xffffu); tile[r][2 * lx + 1] = (unsigned short)(v >> 16);
        }
    }
    __syncthreads();
#pragma unroll 4
    for (int c = ly; c < 64; c += 8) {
        const int gc = tc0 + c, gr = tr + 2 * lx;
        if (gc < cols && gr < rows)
            *reinterpret_cast<uint32_t*>(d + (int64_t)gc * rows + gr) = (uint32_t)tile[2 * lx][c] | ((uint32_t)tile[2 * lx + 1][c] << 16);
    }
}

template <class T>
int layout_transpose_launch(const void* src, void* dst, int64_t n, int rows, int cols, cudaStream_t stream)
{
    if (n == 0 || rows == 0 || cols == 0) return 0;
    if (sizeof(T) == 2 && rows % 2 == 0 && cols % 2 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0) {
        const int tr_ = (rows + 63) / 64, tc_ = (cols + 63) / 64;
        layout_transpose16_kernel<<<dim3((unsigned)(tr_ * tc_), (unsigned)n), 256, 0, stream>>>(static_cast<const unsigned short*>(src),
                                                                                                  static_cast<unsigned short*>(dst), rows, cols, tr_, tc_);
        return check_launch("layout_transpose16_kernel");
    }
    const int tiles_r = (rows + 63) / 64, tiles_c = (cols + 63) / 64;
    const dim3 grid((unsigned)(tiles_r * tiles_c), (unsigned)n);
    layout_transpose_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(src), static_cast<T*>(dst), rows, cols, tiles_r, tiles_c);
    return check_launch("layout_transpose_kernel");
}

// img[n][c][hw] (fp32, contiguous NCHW) += y[n][hw][c] (fp16 or fp32, channels-last): the skip image's `y.to(float32, contiguous_format)`
// and `img.add_(y)` (networks_stylegan2.py:456-457) in one pass over a shared-memory tile — torch runs them as a strided copy (230 us for the
// 96-channel 256^2 image of 8 planes sets) followed by an add (80 us).
template <class T>
__global__ void __launch_bounds__(256) image_accumulate_kernel(const T* __restrict__ src, float* __restrict__ dst, int rows, int cols, int tiles_c)
{
    // src [n][rows = hw][cols = c] -> dst [n][cols][rows]
    __shared__ float tile[64][65];
    const int64_t n = blockIdx.y;
    const int tr = (blockIdx.x / tiles_c) * 64, tc0 = (blockIdx.x % tiles_c) * 64;
    const T* s = src + n * (int64_t)rows * cols;
    float* d = dst + n * (int64_t)rows * cols;
    const int lx = threadIdx.x & 63, ly = threadIdx.x >> 6;
    // (all 16 loads of a thread are issued before the first use: the pass is a read-modify-write and latency-bound otherwise)
    T in[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int gr = tr + ly + 4 * i, gc = tc0 + lx;
        in[i] = (gr < rows && gc < cols) ? s[(int64_t)gr * cols + gc] : T(0.0f);
    }
    float old[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int gc = tc0 + ly + 4 * i, gr = tr + lx;
        old[i] = (gc < cols && gr < rows) ? d[(int64_t)gc * rows + gr] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[ly + 4 * i][lx] = (float)in[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = ly + 4 * i, gc = tc0 + c, gr = tr + lx;
        if (gc < cols && gr < rows) d[(int64_t)gc * rows + gr] = old[i] + tile[lx][c];
    }
}

// a handful of channels (RGB): one thread per pixel, plane writes coalesced
template <class T>
__global__ void __launch_bounds__(256) image_accumulate_small_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t hw, int c)
{
    const int64_t n = blockIdx.y;
    const T* s = src + n * hw * c;
    float* d = dst + n * hw * c;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < hw; p += (int64_t)gridDim.x * 256)
        for (int k = 0; k < c; ++k) d[k * hw + p] += (float)s[p * c + k];
}

template <class T>
static int image_accumulate_launch(const void* src, float* dst, int64_t n, int c, int64_t hw, cudaStream_t stream)
{
    if (c <= 4) {
        const dim3 grid((unsigned)std::min<int64_t>((hw + 255) / 256, 4096), (unsigned)n);
        image_accumulate_small_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(src), dst, hw, c);
        return check_launch("image_accumulate_small_kernel");
    }
    const int tiles_r = (int)((hw + 63) / 64), tiles_c = (c + 63) / 64;
    const dim3 grid((unsigned)(tiles_r * tiles_c), (unsigned)n);
    image_accumulate_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(src), dst, (int)hw, c, tiles_c);
    return check_launch("image_accumulate_kernel");
}

}  // namespace sg
}  // namespace nfe

using namespace nfe;

NFE_EXPORT int nfe_image_accumulate(const void* y_channels_last, float* img, int64_t n, int c, int64_t hw, int dtype, nfe_stream_t stream)
{
    NFE_REQUIRE(n >= 0 && c >= 0 && hw >= 0 && hw < (1ll << 31) && n <= 65535, "nfe_image_accumulate: bad shape");
    if (n * c * hw == 0) return 0;
    NFE_REQUIRE(y_channels_last && img, "nfe_image_accumulate: null pointer");
    if (dtype == NFE_DTYPE_F32) return sg::image_accumulate_launch<float>(y_channels_last, img, n, c, hw, as_stream(stream));
    if (dtype == NFE_DTYPE_F16) return sg::image_accumulate_launch<__half>(y_channels_last, img, n, c, hw, as_stream(stream));
    set_error("nfe_image_accumulate: dtype must be NFE_DTYPE_F32 / F16, got %d", dtype);
    return 1;
}

NFE_EXPORT int nfe_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y, int64_t size_x,
                            int size_b, int64_t step_b, int dtype, int grad, int act, float alpha, float gain, float clamp, nfe_stream_t stream)
{
    NFE_REQUIRE(size_x >= 0 && (size_x == 0 || (x && y)), "nfe_bias_act: null x / y");
    NFE_REQUIRE(grad >= 0 && grad <= 2, "nfe_bias_act: grad must be 0, 1 or 2, got %d", grad);
    NFE_REQUIRE(!b || (size_b > 0 && step_b > 0), "nfe_bias_act: bias needs size_b > 0 and step_b > 0");
    NFE_REQUIRE(grad == 0 || yref || act == 1 || act == 9, "nfe_bias_act: grad >= 1 needs the saved output (yref)");
    NFE_REQUIRE(!(act == 9 && grad > 0) || xref, "nfe_bias_act: swish gradients need the saved input (xref)");
    NFE_REQUIRE(grad < 2 || dy, "nfe_bias_act: grad == 2 needs dy");
    if (size_x == 0) return 0;
    sg::BiasActArgs p{x, b, xref, yref, dy, y, size_x, b ? size_b : 1, b ? step_b : 1, grad, alpha, gain, clamp};
    if (dtype == NFE_DTYPE_F32) return sg::bias_act_dispatch<float>(act, p, as_stream(stream));
    if (dtype == NFE_DTYPE_F16) return sg::bias_act_dispatch<__half>(act, p, as_stream(stream));
    if (dtype == NFE_DTYPE_BF16) return sg::bias_act_dispatch<__nv_bfloat16>(act, p, as_stream(stream));
    set_error("nfe_bias_act: dtype must be NFE_DTYPE_F32 / F16 / BF16, got %d", dtype);
    return 1;
}

NFE_EXPORT int nfe_upfirdn2d(const void* x, const float* f, void* y, int n, int c, int in_h, int in_w, int out_h, int out_w, int fh, int fw,
                             const int64_t* x_strides, const int64_t* y_strides, int upx, int upy, int downx, int downy, int padx0,
                             int padx1, int pady0, int pady1, int flip_filter, float gain, int dtype, nfe_stream_t stream)
{
    NFE_REQUIRE(x && f && y && x_strides && y_strides, "nfe_upfirdn2d: null pointer");
    NFE_REQUIRE(n >= 0 && c >= 0 && in_h > 0 && in_w > 0 && fh >= 1 && fw >= 1, "nfe_upfirdn2d: bad shape");
    NFE_REQUIRE(upx >= 1 && upy >= 1 && downx >= 1 && downy >= 1, "nfe_upfirdn2d: scaling factors must be >= 1");
    // upfirdn2d.cpp:55-58 / upfirdn2d.py:185-187: the up-sampled, padded image must not be smaller than the filter
    const int64_t up_w = (int64_t)in_w * upx + padx0 + padx1, up_h = (int64_t)in_h * upy + pady0 + pady1;
    NFE_REQUIRE(up_w >= fw && up_h >= fh, "nfe_upfirdn2d: up-sampled image (%lld x %lld) smaller than the filter (%d x %d)", (long long)up_h,
                (long long)up_w, fh, fw);
    NFE_REQUIRE(out_w == (up_w - fw + downx) / downx && out_h == (up_h - fh + downy) / downy, "nfe_upfirdn2d: output must be %lld x %lld, got %d x %d",
                (long long)((up_h - fh + downy) / downy), (long long)((up_w - fw + downx) / downx), out_h, out_w);
    sg::UpfirdnArgs p;
    p.x = x; p.f = f; p.y = y; p.n = n; p.c = c; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.fh = fh; p.fw = fw;
    for (int i = 0; i < 4; ++i) { p.xs[i] = x_strides[i]; p.ys[i] = y_strides[i]; }
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip_filter ? 1 : 0; p.gain = gain;
    if (dtype == NFE_DTYPE_F32) return sg::upfirdn2d_launch<float>(p, as_stream(stream));
    if (dtype == NFE_DTYPE_F16) return sg::upfirdn2d_launch<__half>(p, as_stream(stream));
    if (dtype == NFE_DTYPE_BF16) return sg::upfirdn2d_launch<__nv_bfloat16>(p, as_stream(stream));
    set_error("nfe_upfirdn2d: dtype must be NFE_DTYPE_F32 / F16 / BF16, got %d", dtype);
    return 1;
}

NFE_EXPORT int nfe_layout_convert(const void* src, void* dst, int64_t n, int c, int64_t hw, int dtype, int to_channels_last, nfe_stream_t stream)
{
    NFE_REQUIRE(n >= 0 && c >= 0 && hw >= 0 && hw < (1ll << 31) && n <= 65535, "nfe_layout_convert: bad shape");
    NFE_REQUIRE((src && dst) || n * c * hw == 0, "nfe_layout_convert: null pointer");
    // NCHW -> channels-last transposes [c][hw] to [hw][c]; the other direction [hw][c] to [c][hw]
    const int rows = to_channels_last ? c : (int)hw, cols = to_channels_last ? (int)hw : c;
    if (dtype == NFE_DTYPE_F32) return sg::layout_transpose_launch<float>(src, dst, n, rows, cols, as_stream(stream));
    if (dtype == NFE_DTYPE_F16 || dtype == NFE_DTYPE_BF16) return sg::layout_transpose_launch<unsigned short>(src, dst, n, rows, cols, as_stream(stream));
    set_error("nfe_layout_convert: dtype must be NFE_DTYPE_F32 / F16 / BF16, got %d", dtype);
    return 1;
}
