// bevipm_shard.cu -- view sharding over peer memory: the fused warp kernel of one rank's cameras ADDS its partial sum
// straight into the BEV row slabs of the ranks that own them (bulk reductions, cp.reduce.async.bulk add.f32, on NVLink-mapped
// peer pointers), so
// the partial BEV never exists in this rank's HBM and the exchange overlaps the warp, tile by tile.
// (BASELINE configs[2] / SURVEY.md 8(e): "views sharded ... with NCCL sum-reduce"; the NCCL forms are in bevipm/sharding.py.)
#include <algorithm>
#include <cstdlib>

#include "../../include/bevipm.h"
#include "ipm_run.cuh"

namespace bevipm {
void note_launch(int variant);
int set_error(int code, const char* msg);
FwdParams params_from_desc(const bevipm_desc* d, const void* feats, const float* K, const float* Rt, const float* xs, const float* ys, void* out);
int check_desc_public(const bevipm_desc* d);

struct FinishArgs {
    const float* buf[16];
    int n;
};

namespace {
template <typename TIn, int MAXREG>
int launch_red(FwdParams p, cudaStream_t st) {
    constexpr int VE = VecTraits<TIn>::VE, CELLS = 8, NW = 4, DEPTH = 4;
    p.tiles_x = (p.Wb + CELLS - 1) / CELLS;
    p.tiles_y = (p.Hb + NW - 1) / NW;
    p.fsy16 = (int)(p.fs_y / VE);
    p.fsx16 = (int)(p.fs_x / VE);
    p.rcpV = 1.0f / (float)p.V;
    auto kern = warp_fuse_run_kernel<TIn, float, CELLS, NW, 1, MAXREG, DEPTH, false, 0, KM_RED>;
    // tables, rings, homographies, (ring barriers), then one parking area of 8 cells x one chunk of fp32 sums per warp
    const size_t smem = ((size_t)run_tables_bytes(p.V, CELLS, NW) + (size_t)NW * DEPTH * 2048 + (size_t)p.V * 48 + (size_t)NW * DEPTH * 8 + 127) / 128 * 128 +
                        (size_t)NW * CELLS * (32 * VE * 4);
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_error(BEVIPM_ERR_CUDA, "cudaFuncSetAttribute (red kernel)");
    int fpc = 1;
    const long long tiles = (long long)p.tiles_x * p.tiles_y, slots = 148LL * (65536 / (MAXREG * 32 * NW));
    while (fpc < 8 && fpc * 2 <= p.B && tiles * ((p.B + fpc * 2 - 1) / (fpc * 2)) >= 16 * slots) fpc *= 2;
    dim3 grid(p.tiles_x * p.tiles_y, 1, (p.B + fpc - 1) / fpc);
    kern<<<grid, NW * 32, smem, st>>>(p, fpc);
    if (cudaGetLastError() != cudaSuccess) return set_error(BEVIPM_ERR_CUDA, "red kernel launch failed");
    note_launch(60);
    return 0;
}
}  // namespace
}  // namespace bevipm

namespace bevipm {
// out[e] = (bufs[0][e] + bufs[1][e] + ...) / divisor, added in the order given (rank order: the result is reproducible),
// IEEE division (fusion.py:20-21); divisor 1 leaves the sum.
__global__ void __launch_bounds__(256) slab_finish_kernel(FinishArgs a, float* __restrict__ out, long long n4, float divisor) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
        float4 s = __ldcs(reinterpret_cast<const float4*>(a.buf[0]) + e);
        for (int q = 1; q < a.n; ++q) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(a.buf[q]) + e);
            s.x = __fadd_rn(s.x, t.x); s.y = __fadd_rn(s.y, t.y); s.z = __fadd_rn(s.z, t.z); s.w = __fadd_rn(s.w, t.w);
        }
        if (divisor != 1.0f) { s.x = __fdiv_rn(s.x, divisor); s.y = __fdiv_rn(s.y, divisor); s.z = __fdiv_rn(s.z, divisor); s.w = __fdiv_rn(s.w, divisor); }
        reinterpret_cast<float4*>(out)[e] = s;
    }
}
}  // namespace bevipm

extern "C" int bevipm_slab_finish(const void* const* bufs, int32_t nbufs, float* out, int64_t n, float divisor, void* stream) {
    using namespace bevipm;
    if (!bufs || !out || nbufs < 1 || nbufs > 16 || n <= 0 || (n & 3)) return set_error(BEVIPM_ERR_BAD_ARG, "slab_finish: 1..16 buffers of n (multiple of 4) floats");
    FinishArgs a;
    a.n = nbufs;
    for (int q = 0; q < nbufs; ++q) {
        if (!bufs[q] || (reinterpret_cast<uintptr_t>(bufs[q]) & 15)) return set_error(BEVIPM_ERR_BAD_ARG, "slab_finish: buffers must be 16-byte aligned");
        a.buf[q] = static_cast<const float*>(bufs[q]);
    }
    const long long n4 = n / 4;
    const long long want = (n4 + 255) / 256;
    slab_finish_kernel<<<(unsigned)std::min<long long>(want, 148LL * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, out, n4, divisor);
    if (cudaGetLastError() != cudaSuccess) return set_error(BEVIPM_ERR_CUDA, "slab_finish launch failed");
    note_launch(61);
    return 0;
}

extern "C" int bevipm_warp_fuse_red(const bevipm_desc* d, const void* feats, const float* K, const float* Rt34, const float* xs,
                                    const float* ys, void* const* slabs, int32_t nslabs, int32_t slab_rows, void* stream) {
    using namespace bevipm;
    if (int rc = check_desc_public(d)) return rc;
    if (!feats || !K || !Rt34 || !xs || !ys || !slabs) return set_error(BEVIPM_ERR_BAD_ARG, "null pointer");
    if (nslabs < 1 || nslabs > 16 || slab_rows < 1 || (long long)nslabs * slab_rows < d->Hb)
        return set_error(BEVIPM_ERR_BAD_ARG, "slabs do not cover the BEV rows (1..16 slabs of slab_rows rows)");
    if (d->mode != BEVIPM_SUM) return set_error(BEVIPM_ERR_UNSUPPORTED, "the peer-memory form adds partial SUMs (divide the owner's slab for mean)");
    if (d->out_dtype != BEVIPM_F32) return set_error(BEVIPM_ERR_UNSUPPORTED, "slabs are f32");
    const int ve = d->in_dtype == BEVIPM_F32 ? 4 : 8;
    const int64_t fs[] = {d->fs_b, d->fs_v, d->fs_y, d->fs_x};
    bool ok = d->fs_c == 1 && d->os_c == 1 && d->C % ve == 0 && d->V <= kRunMaxViews && (reinterpret_cast<uintptr_t>(feats) & 15) == 0;
    for (int64_t s : fs) ok = ok && s % ve == 0 && s >= 0;
    ok = ok && d->os_b % 4 == 0 && d->os_y % 4 == 0 && d->os_x % 4 == 0;
    ok = ok && (long long)d->V * (d->fs_v / ve) + (long long)(d->Hf + 2) * (d->fs_y / ve) + (long long)(d->Wf + 2) * (d->fs_x / ve) <= 0x7fffffffLL;
    if (!ok) return set_error(BEVIPM_ERR_UNSUPPORTED, "the peer-memory form needs channels-last features and slabs (16-byte vectors), V <= 32");
    FwdParams p = params_from_desc(d, feats, K, Rt34, xs, ys, nullptr);
    for (int q = 0; q < nslabs; ++q) {
        if (!slabs[q] || (reinterpret_cast<uintptr_t>(slabs[q]) & 15)) return set_error(BEVIPM_ERR_BAD_ARG, "slab pointers must be 16-byte aligned");
        p.slab[q] = slabs[q];
    }
    p.slab_rows = slab_rows;
    p.slab_put = (d->flags & BEVIPM_FLAG_SLAB_PUT) ? 1 : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return d->in_dtype == BEVIPM_F32 ? launch_red<float, 96>(p, st) : launch_red<__nv_bfloat16, 128>(p, st);
}
