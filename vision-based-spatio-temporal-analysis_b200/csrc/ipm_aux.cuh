// ipm_aux.cuh -- the kernels around the fused fast path: strided fallback forward, backward,
// sample-coordinate dump, NCHW -> NHWC layout pre-pass.  All CUDA; there is no CPU path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ipm_fused.cuh"

namespace bevipm {

template <typename T> __device__ __forceinline__ float load_f32(const T* p);
template <> __device__ __forceinline__ float load_f32<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
}
template <typename T> __device__ __forceinline__ void store_f32(T* p, float v);
template <> __device__ __forceinline__ void store_f32<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_f32<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
}

constexpr int kGenTH = 4, kGenTW = 32;  // strided kernels: one thread per cell of a 4 x 32 patch

// One view's bilinear value for one channel: fma(SE,se, fma(SW,sw, fma(NE,ne, NW*nw))).
template <typename TIn>
__device__ __forceinline__ float sample_scalar(const TIn* plane, const CellTap& t, long long fs_y, long long fs_x) {
    const TIn* base = plane + (long long)t.y0 * fs_y + (long long)t.x0 * fs_x;
    const float a = (t.flags & 1) ? load_f32(base) : 0.0f;
    const float b = (t.flags & 2) ? load_f32(base + fs_x) : 0.0f;
    const float c = (t.flags & 4) ? load_f32(base + fs_y) : 0.0f;
    const float d = (t.flags & 8) ? load_f32(base + fs_y + fs_x) : 0.0f;
    float r = __fmul_rn(a, t.nw);
    r = __fmaf_rn(b, t.ne, r);
    r = __fmaf_rn(c, t.sw, r);
    r = __fmaf_rn(d, t.se, r);
    return r;
}

// Any strides, any C (the NCHW-contiguous tensors the reference's encoder emits, odd channel
// counts).  Lanes run along BEV x so NCHW stores coalesce; channels are looped per thread.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kGenTH* kGenTW) warp_fuse_strided_kernel(const FwdParams p, int c_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellTap* taps = reinterpret_cast<CellTap*>(smem_raw);
    float* sH = reinterpret_cast<float*>(taps + p.V * kGenTH * kGenTW);
    const int b = blockIdx.z;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i0 = ty * kGenTH, j0 = tx * kGenTW;
    project_patch(p, b, i0, j0, kGenTH, kGenTW, taps, sH);
    const int r = threadIdx.x / kGenTW, q = threadIdx.x % kGenTW;
    const int i = i0 + r, j = j0 + q;
    if (i >= p.Hb || j >= p.Wb) return;
    const TIn* fb = reinterpret_cast<const TIn*>(p.feats) + (long long)b * p.fs_b;
    TOut* ob = reinterpret_cast<TOut*>(p.out) + (long long)b * p.os_b + (long long)i * p.os_y + (long long)j * p.os_x;
    const float qnan = __int_as_float(0x7fc00000);
    const float Vf = (float)p.V;
    const int c_end = min(p.C, (int)(blockIdx.y + 1) * c_per_cta);
    for (int c = blockIdx.y * c_per_cta; c < c_end; ++c) {
        float acc = (p.mode == 2) ? -INFINITY : 0.0f;
        for (int v = 0; v < p.V; ++v) {
            const CellTap t = taps[(v * kGenTH + r) * kGenTW + q];
            float s = 0.0f;
            if (t.flags & kTapMask) s = sample_scalar(fb + (long long)v * p.fs_v + (long long)c * p.fs_c, t, p.fs_y, p.fs_x);
            else if (t.flags & kNonFinite) s = qnan;
            if (p.mode == 3) store_f32(ob + (long long)v * p.os_v + (long long)c * p.os_c, s);
            else if (p.mode == 2) acc = (s > acc || s != s) ? s : acc;
            else acc = __fadd_rn(acc, s);
        }
        if (p.mode == 1) acc = __fdiv_rn(acc, Vf);
        if (p.mode != 3) store_f32(ob + (long long)c * p.os_c, acc);
    }
}

// Backward w.r.t. the features: scatter of the same taps (autograd of geometry.py:161 followed
// by fusion.py:18-21).  grad_feats is fp32 and pre-zeroed; out-strides describe grad_out.
// VEC: channels-last on both sides, C % 4 == 0 -> lane owns 4 channels, red.global.add.v4.f32.
template <typename TG, bool VEC>
__global__ void __launch_bounds__(kGenTH* kGenTW) warp_fuse_bwd_kernel(const FwdParams p, int c_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellTap* taps = reinterpret_cast<CellTap*>(smem_raw);
    float* sH = reinterpret_cast<float*>(taps + p.V * kGenTH * kGenTW);
    const int b = blockIdx.z;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i0 = ty * kGenTH, j0 = tx * kGenTW;
    project_patch(p, b, i0, j0, kGenTH, kGenTW, taps, sH);
    const TG* gb = reinterpret_cast<const TG*>(p.out) + (long long)b * p.os_b;
    float* fb = reinterpret_cast<float*>(const_cast<void*>(p.feats)) + (long long)b * p.fs_b;
    const float Vf = (float)p.V;
    const int c_begin = blockIdx.y * c_per_cta, c_end = min(p.C, c_begin + c_per_cta);
    if constexpr (VEC) {
        // warp r walks the 32 cells of its row; lanes own 4 consecutive channels each
        const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int i = i0 + r;
        if (i >= p.Hb) return;
        for (int q = 0; q < kGenTW; ++q) {
            const int j = j0 + q;
            if (j >= p.Wb) break;
            const TG* gc = gb + (long long)i * p.os_y + (long long)j * p.os_x;
            for (int c = c_begin + lane * 4; c < c_end; c += 128) {
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.mode != 3) {
                    g = make_float4(load_f32(gc + c), load_f32(gc + c + 1), load_f32(gc + c + 2), load_f32(gc + c + 3));
                    if (p.mode == 1) g = make_float4(__fdiv_rn(g.x, Vf), __fdiv_rn(g.y, Vf), __fdiv_rn(g.z, Vf), __fdiv_rn(g.w, Vf));
                }
                for (int v = 0; v < p.V; ++v) {
                    const CellTap t = taps[(v * kGenTH + r) * kGenTW + q];
                    if (!(t.flags & kTapMask)) continue;
                    if (p.mode == 3) {
                        const TG* gv = gc + (long long)v * p.os_v + c;
                        g = make_float4(load_f32(gv), load_f32(gv + 1), load_f32(gv + 2), load_f32(gv + 3));
                    }
                    float* base = fb + (long long)v * p.fs_v + (long long)t.y0 * p.fs_y + (long long)t.x0 * p.fs_x + c;
                    const float w[4] = {t.nw, t.ne, t.sw, t.se};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (!((t.flags >> k) & 1)) continue;
                        float* tp = base + (k & 1 ? p.fs_x : 0) + (k & 2 ? p.fs_y : 0);
                        atomicAdd(reinterpret_cast<float4*>(tp),
                                  make_float4(__fmul_rn(w[k], g.x), __fmul_rn(w[k], g.y), __fmul_rn(w[k], g.z), __fmul_rn(w[k], g.w)));
                    }
                }
            }
        }
    } else {
        const int r = threadIdx.x / kGenTW, q = threadIdx.x % kGenTW;
        const int i = i0 + r, j = j0 + q;
        if (i >= p.Hb || j >= p.Wb) return;
        const TG* gc = gb + (long long)i * p.os_y + (long long)j * p.os_x;
        for (int c = c_begin; c < c_end; ++c) {
            float g = 0.0f;
            if (p.mode != 3) {
                g = load_f32(gc + (long long)c * p.os_c);
                if (p.mode == 1) g = __fdiv_rn(g, Vf);
            }
            for (int v = 0; v < p.V; ++v) {
                const CellTap t = taps[(v * kGenTH + r) * kGenTW + q];
                if (!(t.flags & kTapMask)) continue;
                if (p.mode == 3) g = load_f32(gc + (long long)v * p.os_v + (long long)c * p.os_c);
                float* base = fb + (long long)v * p.fs_v + (long long)c * p.fs_c + (long long)t.y0 * p.fs_y + (long long)t.x0 * p.fs_x;
                if (t.flags & 1) atomicAdd(base, __fmul_rn(t.nw, g));
                if (t.flags & 2) atomicAdd(base + p.fs_x, __fmul_rn(t.ne, g));
                if (t.flags & 4) atomicAdd(base + p.fs_y, __fmul_rn(t.sw, g));
                if (t.flags & 8) atomicAdd(base + p.fs_y + p.fs_x, __fmul_rn(t.se, g));
            }
        }
    }
}

// ix, iy [B,V,Hb,Wb]: what phase A computes, exported for tests and for the byte counter.
__global__ void sample_coords_kernel(const FwdParams p, float* __restrict__ ix, float* __restrict__ iy) {
    const int bv = blockIdx.y;
    __shared__ float H[9];
    if (threadIdx.x == 0) homography(p.K + 9 * bv, p.Rt + 12 * bv, H);
    __syncthreads();
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= p.Hb * p.Wb) return;
    const int i = cell / p.Wb, j = cell - i * p.Wb;
    float x, y;
    cell_coord(H, p.xs[j], p.ys[i], p.sw, p.sh, (float)p.Wf, (float)p.Hf, x, y, p.kx, p.ky);
    ix[(long long)bv * p.Hb * p.Wb + cell] = x;
    iy[(long long)bv * p.Hb * p.Wb + cell] = y;
}

// Source texels any BEV cell samples, per (frame, view) and source row: span[(bv * Hf + y) * 2] = first,
// [.. + 1] = last x holding an in-map tap (first > last: nothing on that row), and -- when `bits` is given -- a bitmap of
// exactly the texels that hold one (bits[(bv * Hf + y) * BW + x / 32] bit x % 32; far-field rows are sampled sparsely).
// Same projection and tap validity as the fused kernels; the host-buffer entry uploads only these texels (rows above the
// horizon of a ground-plane homography are never read).  span[] must be preset to {INT_MAX, -1}, bits[] to 0.
// Dynamic shared memory: Hf * (2 + BW) ints.
__global__ void touched_spans_kernel(const FwdParams p, int* __restrict__ span, unsigned* __restrict__ bits, int BW) {
    extern __shared__ int s_span[];  // [Hf][2]: this block's spans, then [Hf][BW] its bitmap; flushed once (few rows per block are touched)
    unsigned* s_bits = reinterpret_cast<unsigned*>(s_span + 2 * p.Hf);
    const int bv = blockIdx.y;
    __shared__ float H[9];
    if (threadIdx.x == 0) homography(p.K + 9 * bv, p.Rt + 12 * bv, H);
    for (int y = threadIdx.x; y < p.Hf; y += blockDim.x) { s_span[2 * y] = 0x7fffffff; s_span[2 * y + 1] = -1; }
    if (bits)
        for (int q = threadIdx.x; q < p.Hf * BW; q += blockDim.x) s_bits[q] = 0u;
    __syncthreads();
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell < p.Hb * p.Wb) {
        const int i = cell / p.Wb, j = cell - i * p.Wb;
        float x, y;
        cell_coord(H, p.xs[j], p.ys[i], p.sw, p.sh, (float)p.Wf, (float)p.Hf, x, y, p.kx, p.ky);
        const CellTap t = make_tap(x, y, p.Wf, p.Hf);
        const int tm = t.flags & kTapMask;
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) {
            if (!((tm >> tap) & 1)) continue;
            const int yy = t.y0 + (tap >> 1), xx = t.x0 + (tap & 1);
            atomicMin(s_span + 2 * yy, xx);
            atomicMax(s_span + 2 * yy + 1, xx);
            if (bits) atomicOr(s_bits + yy * BW + (xx >> 5), 1u << (xx & 31));
        }
    }
    __syncthreads();
    int* sv = span + (long long)bv * p.Hf * 2;
    for (int y = threadIdx.x; y < p.Hf; y += blockDim.x) {
        if (s_span[2 * y] <= s_span[2 * y + 1]) {
            atomicMin(sv + 2 * y, s_span[2 * y]);
            atomicMax(sv + 2 * y + 1, s_span[2 * y + 1]);
        }
    }
    if (bits) {
        unsigned* bvp = bits + (long long)bv * p.Hf * BW;
        for (int q = threadIdx.x; q < p.Hf * BW; q += blockDim.x)
            if (s_bits[q]) atomicOr(bvp + q, s_bits[q]);
    }
}

// Host-buffer entry, upload side: the features lie in PINNED HOST memory that the device can address (UVA); this kernel
// pulls exactly the texels some BEV cell samples (touched_spans_kernel's table: per view and source row [x_lo, x_hi] and the
// bitmap of sampled texels inside it) over PCIe into the device arena.  A warp takes one sampled texel at a time, 16 bytes per
// lane and load, up to four loads in flight per lane.  One launch per frame replaces ~60 banded cudaMemcpy2DAsync calls, moves no
// slack bytes and needs no read-back of the table to the host.  blockIdx.y = (view, source row); gridDim.x CTAs share a row.
__global__ void __launch_bounds__(256) host_span_gather_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int* __restrict__ spans,
                                                                const unsigned* __restrict__ bits, int BW, int Wf, int texel16,
                                                                const int* __restrict__ dma_rows, int Hf) {
    const int row = blockIdx.y;
    {   // rows the copy engine uploads (a run of dense rows per view; empty run: first > last)
        const int v = row / Hf, y = row - v * Hf;
        if (y >= dma_rows[2 * v] && y <= dma_rows[2 * v + 1]) return;
    }
    const int lo = spans[2 * row], hi = spans[2 * row + 1];
    if (lo > hi) return;
    const unsigned* rb = bits + (long long)row * BW;
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    for (int t = lo + blockIdx.x * wpc + (threadIdx.x >> 5); t <= hi; t += gridDim.x * wpc) {
        if (!((rb[t >> 5] >> (t & 31)) & 1u)) continue;   // inside the span, but no cell samples it (warp-uniform)
        const long long base = ((long long)row * Wf + t) * texel16;
        const uint4* s = src + base;
        uint4* d = dst + base;
        int e = lane;
        for (; e + 96 < texel16; e += 128) {
            uint4 a0, a1, a2, a3;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w) : "l"(s + e));
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a1.x), "=r"(a1.y), "=r"(a1.z), "=r"(a1.w) : "l"(s + e + 32));
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a2.x), "=r"(a2.y), "=r"(a2.z), "=r"(a2.w) : "l"(s + e + 64));
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a3.x), "=r"(a3.y), "=r"(a3.z), "=r"(a3.w) : "l"(s + e + 96));
            d[e] = a0; d[e + 32] = a1; d[e + 64] = a2; d[e + 96] = a3;
        }
        for (; e < texel16; e += 32) d[e] = __ldg(s + e);
    }
}

// fusion.py:17-22 on materialised maps: in [B,V,inner] -> out [B,inner]; sequential over v.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) fuse_views_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int V,
                                                         long long inner, int mode) {
    const long long b = blockIdx.y;
    const TIn* ib = in + b * V * inner;
    TOut* ob = out + b * inner;
    const float Vf = (float)V;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < inner; e += (long long)gridDim.x * blockDim.x) {
        float acc = load_f32(ib + e);
        for (int v = 1; v < V; ++v) {
            const float s = load_f32(ib + (long long)v * inner + e);
            if (mode == 2) acc = (s > acc || s != s) ? s : acc;
            else acc = __fadd_rn(acc, s);
        }
        if (mode == 1) acc = __fdiv_rn(acc, Vf);
        store_f32(ob + e, acc);
    }
}

// Backward of fuse_views_kernel: grad_in[b,v,e] from grad_out[b,e].  sum: g; mean: g / V; max: g to the FIRST view that
// holds the maximum (torch.max's index rule), 0 to the others; a NaN maximum sends g to the first NaN view.
template <typename TIn>
__global__ void __launch_bounds__(256) fuse_views_bwd_kernel(const TIn* __restrict__ in, const float* __restrict__ gout,
                                                             float* __restrict__ gin, int V, long long inner, int mode) {
    const long long b = blockIdx.y;
    const TIn* ib = in + b * V * inner;
    float* gb = gin + b * V * inner;
    const float Vf = (float)V;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < inner; e += (long long)gridDim.x * blockDim.x) {
        const float g = gout[b * inner + e];
        if (mode == 2) {
            float best = load_f32(ib + e);
            int arg = 0;
            for (int v = 1; v < V; ++v) {
                const float s = load_f32(ib + (long long)v * inner + e);
                if (best == best && (s > best || s != s)) { best = s; arg = v; }
            }
            for (int v = 0; v < V; ++v) gb[(long long)v * inner + e] = v == arg ? g : 0.0f;
        } else {
            const float gv = mode == 1 ? __fdiv_rn(g, Vf) : g;
            for (int v = 0; v < V; ++v) gb[(long long)v * inner + e] = gv;
        }
    }
}

// Validity mask count (north star: "validity-mask counts"): count[b,i,j] = number of views whose sample position of BEV
// cell (i, j) has at least one bilinear tap inside the feature map (the cells geometry.py:161 reads non-padding from).
// One warp per 32 cells of a row, views looped; same projection and tap test as the fused kernels.
__global__ void __launch_bounds__(256) valid_count_kernel(const FwdParams p, int* __restrict__ count) {
    extern __shared__ float s_h[];  // [V][9]
    const int b = blockIdx.y;
    for (int v = threadIdx.x; v < p.V; v += blockDim.x) homography(p.K + 9 * (b * p.V + v), p.Rt + 12 * (b * p.V + v), s_h + 9 * v);
    __syncthreads();
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= p.Hb * p.Wb) return;
    const int i = cell / p.Wb, j = cell - i * p.Wb;
    const float x = p.xs[j], y = p.ys[i];
    int n = 0;
    for (int v = 0; v < p.V; ++v) {
        float ix, iy;
        cell_coord(s_h + 9 * v, x, y, p.sw, p.sh, (float)p.Wf, (float)p.Hf, ix, iy, p.kx, p.ky);
        n += (make_tap(ix, iy, p.Wf, p.Hf).flags & kTapMask) != 0;
    }
    count[(long long)b * p.Hb * p.Wb + cell] = n;
}

// Mean over the views that SEE a cell (opt-in extension; the reference's mean divides by V, fusion.py:20-21):
// bev[b,i,j,c] (a SUM-mode result, fp32, element strides os_*) /= max(count[b,i,j], 1), IEEE division.
__global__ void __launch_bounds__(128) divide_by_count_kernel(float* __restrict__ bev, const int* __restrict__ count, int C, int Hb, int Wb,
                                                              long long os_b, long long os_c, long long os_y, long long os_x) {
    const long long bc = blockIdx.x;  // (frame, cell)
    const int cells = Hb * Wb;
    const int b = (int)(bc / cells), cell = (int)(bc - (long long)b * cells);
    const int n = count[bc];
    if (n <= 1) return;
    const int i = cell / Wb, j = cell - i * Wb;
    const float nf = (float)n;
    float* o = bev + (long long)b * os_b + (long long)i * os_y + (long long)j * os_x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) o[(long long)c * os_c] = __fdiv_rn(o[(long long)c * os_c], nf);
}

// [N,C,HW] -> [N,HW,C]   (the encoder's NCHW maps, cnn_encoder.py:65-70, into the fast path's layout)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int HW) {
    __shared__ T tile[32][33];
    const long long n = blockIdx.z;
    const int c0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T* sp = src + n * (long long)C * HW;
    T* dp = dst + n * (long long)C * HW;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, s = s0 + tx;
        if (c < C && s < HW) tile[ty + 8 * k][tx] = sp[(long long)c * HW + s];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = s0 + ty + 8 * k, c = c0 + tx;
        if (c < C && s < HW) dp[(long long)s * C + c] = tile[tx][ty + 8 * k];
    }
}

}  // namespace bevipm
