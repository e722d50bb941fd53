/*
 * bevipm.h -- C ABI of the B200 (sm_100a) multi-view IPM warp + BEV fusion library.
 *
 * The reference (sea-sky-web/Vision-based-Spatio-Temporal-Analysis) is pure Python and has no
 * native boundary of its own; the interface this ABI stands behind is the pair of torch modules
 *
 *   GeometryTransformer.forward(feats, intrinsics, extrinsics, img_size) -> [B,V,C,Hb,Wb]
 *       /root/reference/project/models/fusion/geometry.py:80-163   (grid_sample branch :142-162)
 *   SimpleFusion.forward / ConcatFusion.forward (bev_maps) -> [B,C,Hb,Wb] / [B,V*C,Hb,Wb]
 *       /root/reference/project/models/fusion/fusion.py:11-22, :39-46
 *
 * called back to back from BEVNet.forward (models/model_wrapper.py:68-69).  Each entry point
 * below names the reference lines it replaces.  The Python binding a maintainer adds is the
 * ctypes stub in INTEGRATION.md (bevipm/_lib.py is that stub, shipped).
 *
 * Conventions
 *   - plain C types only; every pointer is caller-owned; the library never allocates, frees or
 *     retains device memory (bevipm_warp_fuse_host is the one exception: it keeps a per-thread
 *     device staging arena, released by bevipm_host_release).
 *   - device entry points are stream-ordered on `stream` (a cudaStream_t passed as void*); they
 *     never synchronise the device.
 *   - return value: 0 on success, negative bevipm_status on failure; bevipm_last_error() gives
 *     the thread-local message.  Nothing throws across this boundary.
 *   - all strides are in ELEMENTS of the tensor's dtype.
 */
#ifndef BEVIPM_H_
#define BEVIPM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEVIPM_VERSION 200 /* 0.2.0 */

enum bevipm_status {
    BEVIPM_OK = 0,
    BEVIPM_ERR_BAD_ARG = -1,     /* null pointer, non-positive extent, unknown enum */
    BEVIPM_ERR_UNSUPPORTED = -2, /* legal request this build has no kernel for */
    BEVIPM_ERR_CUDA = -3         /* a CUDA runtime call failed; message has cudaGetErrorString */
};

enum bevipm_dtype { BEVIPM_F32 = 0, BEVIPM_BF16 = 1 };

/* View-axis reduction.  SUM/MEAN/MAX = SimpleFusion modes (fusion.py:17-22; MEAN divides by V,
 * MAX competes against the zeros of out-of-view cells).  NONE writes the per-view maps
 * [B,V,C,Hb,Wb], i.e. GeometryTransformer's own output, which ConcatFusion merely reshapes. */
enum bevipm_mode { BEVIPM_SUM = 0, BEVIPM_MEAN = 1, BEVIPM_MAX = 2, BEVIPM_NONE = 3 };

/* bevipm_desc.flags: sample positions of the reference's kornia branch (geometry.py:124-141: kornia's
 * warp_perspective normalises pixel grids with (size-1) and samples with align_corners=False, so a source pixel p is
 * read at p*size/(size-1) - 0.5; the caller passes BEV cell CORNERS x_min + j*res_x in xs/ys, which is where that
 * branch puts BEV pixel j).  From kornia's published algorithm; kornia is not installed here: parity unpinned. */
#define BEVIPM_FLAG_KORNIA_GEOMETRY 1
/* bevipm_warp_fuse_red only: store the partial sums instead of adding them (every rank owns a receive buffer at each
 * owner; the owner sums them with bevipm_slab_finish): no atomics, no zeroing, reproducible sum order. */
#define BEVIPM_FLAG_SLAB_PUT 2

typedef struct bevipm_desc {
    int32_t B, V, C;        /* frames, views (cameras), channels */
    int32_t Hf, Wf;         /* feature-map size */
    int32_t Hb, Wb;         /* BEV grid size */
    int32_t img_h, img_w;   /* the img_size argument of GeometryTransformer.forward (geometry.py:83) */
    int32_t mode;           /* bevipm_mode */
    int32_t in_dtype;       /* bevipm_dtype of feats (fwd) / grad_feats is always f32 (bwd) */
    int32_t out_dtype;      /* bevipm_dtype of out (fwd) / grad_out (bwd) */
    int32_t variant;        /* 0 = library picks the kernel; >0 forces one, for sweeps */
    int32_t flags;          /* BEVIPM_FLAG_* bits; 0 = the reference's grid_sample geometry */
    int64_t fs_b, fs_v, fs_c, fs_y, fs_x; /* feats[b,v,c,y,x] strides; fs_c == 1 is the NHWC fast path */
    int64_t os_b, os_v, os_c, os_y, os_x; /* out[b,(v,)c,i,j] strides; os_v is read only for NONE */
} bevipm_desc;

int bevipm_version(void);
const char *bevipm_last_error(void);

/* Number of kernels this library has launched in the calling process (all threads). */
int64_t bevipm_launch_count(void);

/* Which kernel variant the last bevipm_warp_fuse_fwd call of this thread launched (the `variant` numbers of
 * DESIGN.md; 0 = none yet, -1 = the strided kernel).  bench.py names the dominant kernel from it. */
int32_t bevipm_last_variant(void);

/*
 * Fused forward: replaces geometry.py:120-162 (homography, projection of every BEV cell centre,
 * bilinear zero-padded sampling) and fusion.py:17-22 in ONE launch, without materialising the
 * per-view maps.
 *   feats  device, dtype in_dtype, logical [B,V,C,Hf,Wf] with strides fs_*
 *   K      device f32 [B*V*9]   row-major 3x3 intrinsics           (geometry.py:35-40)
 *   Rt34   device f32 [B*V*12]  row-major [R|t], world -> camera   (geometry.py:41-52)
 *   xs, ys device f32 [Wb], [Hb] cell-centre world coordinates, torch.linspace values (geometry.py:26-27)
 *   out    device, dtype out_dtype, [B,C,Hb,Wb] (or [B,V,C,Hb,Wb] for NONE) with strides os_*
 */
int bevipm_warp_fuse_fwd(const bevipm_desc *d, const void *feats, const float *K, const float *Rt34,
                         const float *xs, const float *ys, void *out, void *stream);

/*
 * Backward w.r.t. the features (autograd of geometry.py:161 + fusion.py:18-21, reached from
 * train.py:243).  grad_feats is f32 with strides fs_*, must be zero-filled by the caller and
 * is accumulated into with atomics.  MAX is not differentiable here (BEVIPM_ERR_UNSUPPORTED).
 */
int bevipm_warp_fuse_bwd(const bevipm_desc *d, const void *grad_out, const float *K, const float *Rt34,
                         const float *xs, const float *ys, float *grad_feats, void *stream);

/* Sample positions only: ix, iy device f32 [B,V,Hb,Wb] in feature pixels (geometry.py:144-158 plus
 * grid_sample's un-normalisation).  Used by tests and by the algorithmic-byte counter. */
int bevipm_sample_coords(const bevipm_desc *d, const float *K, const float *Rt34, const float *xs,
                         const float *ys, float *ix, float *iy, void *stream);

/* Layout pre-pass for callers that hold the encoder's NCHW-contiguous output (cnn_encoder.py:65-70):
 * src [N,C,H,W] -> dst [N,H,W,C], same dtype (bevipm_dtype). */
int bevipm_nchw_to_nhwc(const void *src, void *dst, int32_t N, int32_t C, int32_t H, int32_t W,
                        int32_t dtype, void *stream);

/* SimpleFusion on already materialised per-view maps (fusion.py:17-22): in [B,V,inner] contiguous
 * -> out [B,inner], inner = C*Hb*Wb.  mode = SUM / MEAN / MAX.  (The fused forward above never
 * needs this; it exists so the stand-alone SimpleFusion module also runs on our kernels.) */
int bevipm_fuse_views(const void *in, void *out, int64_t B, int32_t V, int64_t inner, int32_t mode,
                      int32_t in_dtype, int32_t out_dtype, void *stream);

/* Backward of bevipm_fuse_views (autograd of fusion.py:17-22): grad_in [B,V,inner] f32 from grad_out [B,inner] f32 and, for
 * MAX, the forward input `in` [B,V,inner] (in_dtype): the gradient goes to the first view holding the maximum, as torch.max
 * does.  `in` may be null for SUM / MEAN. */
int bevipm_fuse_views_bwd(const void *in, const float *grad_out, float *grad_in, int64_t B, int32_t V, int64_t inner,
                          int32_t mode, int32_t in_dtype, void *stream);

/*
 * Validity-mask counts (north star): count [B,Hb,Wb] int32 = number of views that see each BEV cell, i.e. whose sample
 * position has at least one bilinear tap inside the feature map -- the cells where geometry.py:161 reads anything but
 * zero padding.  Uses d's extents, img size and flags; strides and dtypes are ignored.  The reference has no such output
 * (its mean divides by V, fusion.py:20-21): an extension, off unless asked for, computed by its own small launch.
 */
int bevipm_valid_count(const bevipm_desc *d, const float *K, const float *Rt34, const float *xs, const float *ys,
                       int32_t *count, void *stream);

/* Opt-in "mean over the views that see the cell": divides a SUM-mode f32 result `bev` (strides d->os_*) in place by
 * max(count, 1), IEEE division.  NOT the reference's mean (that is BEVIPM_MEAN: / V). */
int bevipm_divide_by_count(const bevipm_desc *d, float *bev, const int32_t *count, void *stream);

/*
 * View sharding over peer memory (BASELINE configs[2]): the fused warp of THIS rank's cameras (feats / K /
 * Rt34 hold only them: d->V = the rank's view count) adds its partial SUM into fp32 BEV row slabs owned by the ranks:
 * BEV rows [q*slab_rows, (q+1)*slab_rows) go to slabs[q] (q < nslabs <= 16), each a buffer [B, slab_rows, Wb, C] with element
 * strides d->os_b / os_y / os_x / os_c == 1 -- the rank's own memory or a peer's, mapped over NVLink (CUDA IPC, symmetric
 * memory).  16-byte red.global.add: the partial never lands in this rank's HBM.  Callers zero the slabs and barrier
 * across ranks before and after (bevipm/sharding.py: PeerSlabFusion).  d->mode must be BEVIPM_SUM, out_dtype f32.
 * The sums leave the kernel as bulk copies from shared memory (cp.reduce.async.bulk add.f32, or plain cp.async.bulk stores
 * with BEVIPM_FLAG_SLAB_PUT, when slabs[q] is THIS rank's private receive buffer at owner q).
 */
int bevipm_warp_fuse_red(const bevipm_desc *d, const void *feats, const float *K, const float *Rt34, const float *xs,
                         const float *ys, void *const *slabs, int32_t nslabs, int32_t slab_rows, void *stream);

/* Owner side of the PUT form: out[e] = (bufs[0][e] + ... + bufs[nbufs-1][e]) / divisor over n floats (n % 4 == 0), added in
 * the order given, IEEE division; divisor 1 = plain sum. */
int bevipm_slab_finish(const void *const *bufs, int32_t nbufs, float *out, int64_t n, float divisor, void *stream);

/*
 * Phase-2 follow-on: multi-view, multi-head deformable-attention sampling (the slot the reference's
 * AttentionFusion placeholder reserves, fusion.py:25-36; semantics = Deformable-DETR's MSDeformAttn with
 * one level per camera view, see csrc/deform_attn.cuh).  All device pointers:
 *   value  [B,S,M,D] value_dtype, S = sum_l H_l*W_l     shapes [L,2] int32 (H,W)   level_start [L] int64
 *   loc    [B,Q,M,L,P,2] f32, (x,y) in [0,1]            attn   [B,Q,M,L,P] f32      out [B,Q,M*D] out_dtype
 * D * sizeof(value element) must be a multiple of 16 and at most 512 bytes.
 */
typedef struct bevipm_deform_desc {
    int32_t B, Q, M, D, L, P;
    int32_t value_dtype, out_dtype; /* bevipm_dtype */
    int64_t S;
} bevipm_deform_desc;

int bevipm_deform_attn_fwd(const bevipm_deform_desc *d, const void *value, const int32_t *shapes,
                           const int64_t *level_start, const float *loc, const float *attn, void *out,
                           void *stream);

/* Backward of bevipm_deform_attn_fwd: grad_value [B,S,M,D] f32 (PRE-ZEROED by the caller, accumulated with atomics), grad_loc
 * [B,Q,M,L,P,2] f32 and grad_attn [B,Q,M,L,P] f32 from grad_out [B,Q,M*D] (d->out_dtype).  Any of the three may be null. */
int bevipm_deform_attn_bwd(const bevipm_deform_desc *d, const void *value, const int32_t *shapes,
                           const int64_t *level_start, const float *loc, const float *attn, const void *grad_out,
                           float *grad_value, float *grad_loc, float *grad_attn, void *stream);

/*
 * Table cache across launches ("cache per-camera coefficient tables / precomputed tap indices + weights";
 * the `_grid_cache` the reference declares and never fills, geometry.py:22).  Cameras are static (wildtrack_loader.py:291-293
 * reads one calibration per camera), so everything the fused kernel derives from the calibration -- per row segment: the blend
 * weights of every (view, cell), the tap offsets of every 2x2 block the row enters, the views that see it -- can be kept on the
 * device between calls.  `plan` is a caller-owned device buffer of bevipm_plan_bytes(d) bytes, ZEROED before its first use
 * (cudaMemset; zeroing it again re-arms it).  bevipm_warp_fuse_fwd_planned is bevipm_warp_fuse_fwd for launches that take the
 * default run kernel (channels-last, V <= 32, d->variant == 0; anything else: BEVIPM_ERR_UNSUPPORTED, use the plain entry):
 *   - an empty cache is filled by the first launch from ITS FRAME 0 (tables written by the CTAs that build them, the header --
 *     calibration, shapes / strides / axes key -- published by the last of them);
 *   - afterwards every frame whose K and Rt34 equal the cached calibration BIT FOR BIT (compared on the device, per CTA,
 *     together with the ends of xs / ys) copies its tables from the cache instead of computing them; any other frame computes
 *     them as before.  The result is identical to bevipm_warp_fuse_fwd in every case; only the time differs.
 * A cache is bound to the first calibration it sees and to one stream at a time.  Config 2 (one 128-channel chunk per texel,
 * where the table build is 31 % of the kernel's instructions): see DESIGN.md 5.
 */
int64_t bevipm_plan_bytes(const bevipm_desc *d);
int bevipm_warp_fuse_fwd_planned(const bevipm_desc *d, const void *feats, const float *K, const float *Rt34, const float *xs,
                                 const float *ys, void *out, void *plan, int64_t plan_bytes, void *stream);

/*
 * BEVNet's 1x1 projection folded in front of the warp.  The reference concatenates the V warped maps
 * and applies nn.Conv2d(V*C, Co, 1) on the BEV grid (model_wrapper.py:68-73); the warp is linear per channel, so
 *   proj(concat_v(warp_v(f_v))) = sum_v warp_v(W_v f_v) + bias,   W_v = proj.weight[:, v*C:(v+1)*C],
 * and the projection can run on the small source maps: out[m, r, :] = W_(m % V) . x[m, r, :] for every map m = b*V + v and
 * source texel r.  A hand-written tcgen05 GEMM (csrc/bevipm_proj.cu: TMA-fed ring, kind::tf32 MMAs into a TMEM accumulator,
 * TMA-store epilogue), device pointers, fp32:
 *   x    [BV, rows, C]   element strides x_map_stride, x_row_stride, 1   (rows = Hf*Wf of channels-last maps)
 *   w_hi [Co, V, C]      contiguous;  passes = 1: the weights themselves;  passes = 3: their TF32 heads (low 13 mantissa
 *   w_lo [Co, V, C]      bits cleared) and w_lo = w - w_hi, the exact remainders (may be null when passes = 1)
 *   out  [BV, rows, Co]  element strides out_map_stride, out_row_stride, 1
 * passes = 1: one TF32 pass (the arithmetic cuDNN uses for the reference's Conv2d under torch's default allow_tf32);
 * passes = 3: split operands, hi*hi + lo*hi + hi*lo, fp32-grade (~1e-6 relative).  Co: a multiple of 16 in [16, 256];
 * C and all strides multiples of 4; 16-byte aligned pointers.  The same entry computes the input gradient of the
 * projection (x := grad_out, w := the transposed weights).
 */
int bevipm_proj1x1(const float *x, const float *w_hi, const float *w_lo, float *out, int32_t BV, int32_t V, int64_t rows,
                   int32_t C, int32_t Co, int64_t x_row_stride, int64_t x_map_stride, int64_t out_row_stride,
                   int64_t out_map_stride, int32_t passes, void *stream);

/*
 * Host-buffer entry (the call a non-torch integrator makes, and what bench.py's e2e times):
 * feats/out are HOST pointers (pinned for full PCIe rate) laid out as feats [B,V,Hf,Wf,C] and
 * out [B,Hb,Wb,C] ([B,V,Hb,Wb,C] for NONE); K, Rt34, xs, ys are host arrays.  Frames are streamed
 * H2D -> kernel -> D2H through a double-buffered device arena and the call returns after the
 * last byte of `out` has landed.  d's stride fields are ignored.
 */
int bevipm_warp_fuse_host(const bevipm_desc *d, const void *feats, const float *K, const float *Rt34,
                          const float *xs, const float *ys, void *out);
/* The staging arena (two frames of features + BEV on the device, two streams) is PER HOST THREAD and lives until that thread
 * calls bevipm_host_release(): a thread that stops using the entry must call it, nothing frees the arena at thread exit.
 * Pinned, device-addressable `feats` (cudaHostAlloc / torch pin_memory) are pulled over PCIe by a gather kernel that reads
 * exactly the sampled texels (a per-row bitmap, computed once per calibration) while the copy engine uploads each view's run of
 * densely sampled rows beside it; other host memory takes banded 2-D copies of the row spans.  On failure both streams are drained before the
 * call returns, so the caller's buffers are never in flight after it. */
void bevipm_host_release(void);
/* Host-to-device bytes the last bevipm_warp_fuse_host call of this thread copied: the entry uploads, per frame,
 * view and band of 16 source rows, only the span of texels some BEV cell samples (found by one small kernel from
 * the calibration). */
int64_t bevipm_host_last_h2d_bytes(void);

#ifdef __cplusplus
}
#endif
#endif /* BEVIPM_H_ */
