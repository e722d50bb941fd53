"""CPU-side checks: the C-ABI library loads and exports what include/bevipm.h declares (no compute
calls without a GPU), argument errors come back as status codes, and the host logic mirrors the
reference's calibration handling (geometry.py:33-64, :96-118)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "bevipm.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bevipm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bevipm import _lib
    L = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/bevipm.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert L.bevipm_version() == 200


def test_desc_layout_matches_header():
    from bevipm import _lib
    # 14 int32 + 10 int64, naturally aligned: the struct the header declares
    assert ctypes.sizeof(_lib.Desc) == 14 * 4 + 10 * 8
    assert _lib.Desc.fs_b.offset == 56


def test_bad_arguments_return_status_not_crash():
    from bevipm import _lib
    L = _lib.load()
    d = _lib.Desc()
    rc = L.bevipm_warp_fuse_fwd(ctypes.byref(d), None, None, None, None, None, None, None)
    assert rc == -1 and b"non-positive" in L.bevipm_last_error()
    d.B = d.V = d.C = d.Hf = d.Wf = d.Hb = d.Wb = d.img_h = d.img_w = 4
    d.mode = 9
    assert L.bevipm_warp_fuse_fwd(ctypes.byref(d), None, None, None, None, None, None, None) == -1
    d.mode = 1
    d.V = 40
    assert L.bevipm_warp_fuse_fwd(ctypes.byref(d), None, None, None, None, None, None, None) == -2
    d.V = 4
    assert L.bevipm_warp_fuse_fwd(ctypes.byref(d), None, None, None, None, None, None, None) == -1  # null pointers
    assert b"null" in L.bevipm_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(-1)


def test_product_path_rejects_cpu_tensors():
    import bevipm
    geom = bevipm.FusedIPM(8, 8, (-1.0, 1.0, -1.0, 1.0))
    with pytest.raises(RuntimeError, match="CUDA"):
        geom(torch.zeros(1, 1, 4, 4, 4), torch.eye(3), torch.eye(4))
    with pytest.raises((RuntimeError, NotImplementedError)):
        bevipm.ops.warp_fuse(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 3, 4),
                             torch.zeros(8), torch.zeros(8), 8, 8, 1, False, 0)


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "vision-based-spatio-temporal-analysis_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
        assert "ipm_oracle" not in text and "torch_chain" not in text, f


def test_module_surface_matches_reference():
    import bevipm
    g = bevipm.GeometryTransformer(120, 360, (-24, 24, -7.2, 7.2), warp_impl="bogus")
    assert g.warp_impl == "grid_sample"                           # geometry.py:20
    assert g.bev_h == 120 and g.bev_w == 360 and abs(g.res_x - 48 / 360) < 1e-12
    assert tuple(g.ground_grid.shape) == (120, 360, 3)
    assert list(g.state_dict().keys()) == []                      # non-persistent buffer only (geometry.py:21)
    from bevipm import rig
    xs, ys = rig.ground_axes(120, 360, (-24, 24, -7.2, 7.2))
    assert torch.equal(g.ground_grid[0, :, 0], xs) and torch.equal(g.ground_grid[:, 0, 1], ys)
    with pytest.raises(AssertionError):
        bevipm.SimpleFusion("median")                            # fusion.py:14
    x = torch.arange(2 * 3 * 4 * 5 * 6, dtype=torch.float32).reshape(2, 3, 4, 5, 6)
    assert torch.equal(bevipm.ConcatFusion()(x), x.reshape(2, 12, 5, 6))
    with pytest.raises(NotImplementedError):
        bevipm.FusionModule()(x)


def test_homography_helpers_follow_reference():
    import bevipm
    from bevipm import rig
    from oracle import ipm_oracle as orc
    K, Rt = rig.look_at_rig(3, 0)
    GT = bevipm.GeometryTransformer
    for v in range(3):
        H = GT._compute_homography(K[v], Rt[v])
        assert np.array_equal(H.numpy(), orc.homography(K[v].numpy(), Rt[v].numpy()))
        assert torch.equal(H, GT._compute_homography(K[v], Rt[v][:3]))               # 3x4 extrinsics
        Hi = GT._compute_img_to_world_homography(K[v], Rt[v])
        assert torch.allclose(Hi @ H, torch.eye(3), atol=1e-3)
    # fall-backs of geometry.py:35-40, :47-59
    H = GT._compute_homography(torch.eye(2), torch.eye(3))
    assert torch.equal(H, torch.diag(torch.tensor([1000.0, 1000.0, 1.0])) @ torch.tensor([[1.0, 0, 0], [0, 1, 0], [0, 0, 0]]))
    H = GT._compute_homography(torch.eye(3), torch.zeros(5, 5))
    assert torch.equal(H, torch.tensor([[1.0, 0, 0], [0, 1, 0], [0, 0, 0]]))


def test_pack_calibration_accepts_every_reference_form():
    from bevipm import pack_calibration, rig
    B, V = 2, 3
    K, Rt = rig.look_at_rig(V, 0)
    Kb, Rb = K[None].expand(B, -1, -1, -1).contiguous(), Rt[None].expand(B, -1, -1, -1).contiguous()
    dev = torch.device("cpu")
    K0, R0 = pack_calibration(Kb, Rb, B, V, dev)
    assert K0.shape == (B, V, 3, 3) and R0.shape == (B, V, 3, 4) and K0.is_contiguous() and R0.is_contiguous()
    assert torch.equal(R0, Rb[..., :3, :])
    forms = [
        ([[Kb[b, v] for v in range(V)] for b in range(B)], [[Rb[b, v] for v in range(V)] for b in range(B)]),  # nested lists
        (K, Rt),                                   # [V,3,3] / [V,4,4]
        (Kb, Rb[..., :3, :]),                      # 3x4 extrinsics
        ([[Kb[b, v].numpy() for v in range(V)] for b in range(B)], [[Rb[b, v].numpy() for v in range(V)] for b in range(B)]),
    ]
    for ki, ri in forms:
        k, r = pack_calibration(ki, ri, B, V, dev)
        assert torch.equal(k, K0) and torch.equal(r, R0)
    k, r = pack_calibration(K[1], Rt[1], B, V, dev)   # one shared 2-D pair
    assert torch.equal(k, K[1].expand(B, V, -1, -1)) and torch.equal(r, Rt[1][:3].expand(B, V, -1, -1))
    k, r = pack_calibration(torch.stack([K[0], K[1]]), torch.stack([Rt[0], Rt[1]]), 2, 5, dev)  # 3-D, shape[0] != V -> per frame
    assert torch.equal(k[1, 4], K[1]) and torch.equal(r[0, 2], Rt[0][:3])


def test_rig_and_byte_model():
    from bevipm import rig
    from oracle import ipm_oracle as orc
    K, Rt = rig.look_at_rig(7, 0)
    K2, Rt2 = rig.look_at_rig(7, 0)
    assert torch.equal(K, K2) and torch.equal(Rt, Rt2)
    R = Rt[:, :3, :3].double()
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3, dtype=torch.float64).expand(7, -1, -1), atol=1e-5)
    wl = rig.WORKLOADS["c1"]
    xs, ys = rig.ground_axes(*wl.bev_hw, wl.bounds)
    ix, iy = orc.coords(K[None].numpy(), Rt[None].numpy(), xs.numpy(), ys.numpy(), wl.feat_hw, wl.img_size)
    alg = rig.algorithmic_bytes(ix[0], iy[0], wl.feat_hw, wl.channels, 4, 4)
    assert alg["out_bytes"] == 512 * 120 * 360 * 4
    assert alg["b_alg"] < alg["b_full"] and alg["touched_in_bytes"] > 0.2 * (alg["b_full"] - alg["out_bytes"])
    assert 0.5 < alg["taps_in_bounds"] / alg["taps_total"] < 0.95


def test_torch_chain_agrees_with_c_oracle():
    """Two independent restatements (ATen ops vs scalar C) of the same reference lines agree bit for bit."""
    from bevipm import rig
    from oracle import ipm_oracle as orc, torch_chain
    K, Rt = rig.look_at_rig(4, 7)
    xs, ys = rig.ground_axes(24, 64, rig.WILDTRACK_BOUNDS)   # C*Hb*Wb multiple of 64: no ATen sum tail
    f = torch.randn(2, 4, 8, 27, 48, generator=torch.Generator().manual_seed(7))
    Kb, Rb = K[None].expand(2, -1, -1, -1).contiguous(), Rt[None].expand(2, -1, -1, -1).contiguous()
    for mode in ("none", "sum", "mean", "max"):
        a = torch_chain.warp_fuse(f, Kb, Rb, xs, ys, rig.WILDTRACK_IMG_SIZE, mode).numpy()
        b = orc.warp_fuse(f.numpy(), Kb.numpy(), Rb.numpy(), xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, mode)
        assert np.array_equal(a, b), mode


def test_round2_entry_points_check_their_arguments():
    from bevipm import _lib
    L = _lib.load()
    d = _lib.Desc()
    d.B = d.V = d.C = d.Hf = d.Wf = d.Hb = d.Wb = d.img_h = d.img_w = 4
    d.mode = _lib.SUM
    assert L.bevipm_valid_count(ctypes.byref(d), None, None, None, None, None, None) == -1
    assert L.bevipm_divide_by_count(ctypes.byref(d), None, None, None) == -1
    assert L.bevipm_fuse_views_bwd(None, None, None, 1, 2, 8, 0, 0, None) == -1
    assert L.bevipm_warp_fuse_red(ctypes.byref(d), None, None, None, None, None, None, 2, 2, None) == -1
    d.mode = _lib.MEAN   # the peer-memory form adds partial sums only
    dummy = ctypes.c_void_p(16)
    arr = (ctypes.c_void_p * 2)(16, 32)
    assert L.bevipm_warp_fuse_red(ctypes.byref(d), dummy, ctypes.cast(dummy, ctypes.c_void_p), dummy, dummy, dummy, arr, 2, 2, None) == -2


def test_projection_and_table_cache_entry_points_check_their_arguments():
    """bevipm_proj1x1 / bevipm_plan_bytes / bevipm_warp_fuse_fwd_planned: every shape and pointer rule is checked on the host
    before anything touches the device (so these run without a GPU)."""
    from bevipm import _lib
    L = _lib.load()
    p16 = ctypes.c_void_p(1 << 20)
    # null pointers / bad pass count / output-channel rule / 16-byte rules
    assert L.bevipm_proj1x1(None, p16, None, p16, 7, 7, 100, 64, 128, 64, 6400, 128, 12800, 1, None) == -1
    assert L.bevipm_proj1x1(p16, p16, None, p16, 7, 7, 100, 64, 128, 64, 6400, 128, 12800, 2, None) == -1
    assert L.bevipm_proj1x1(p16, p16, None, p16, 7, 7, 100, 64, 128, 64, 6400, 128, 12800, 3, None) == -1      # split mode without w_lo
    assert L.bevipm_proj1x1(p16, p16, None, p16, 7, 7, 100, 64, 24, 64, 6400, 24, 2400, 1, None) == -2          # Co not a multiple of 16
    assert L.bevipm_proj1x1(p16, p16, None, p16, 7, 7, 100, 64, 272, 64, 6400, 272, 27200, 1, None) == -2       # Co > 256
    assert L.bevipm_proj1x1(p16, p16, None, p16, 7, 7, 100, 30, 128, 30, 3000, 128, 12800, 1, None) == -2       # C not a multiple of 4
    assert L.bevipm_proj1x1(ctypes.c_void_p((1 << 20) + 4), p16, None, p16, 7, 7, 100, 64, 128, 64, 6400, 128, 12800, 1, None) == -1
    assert b"proj1x1" in L.bevipm_last_error()
    d = _lib.Desc()
    d.B, d.V, d.C, d.Hf, d.Wf, d.Hb, d.Wb, d.img_h, d.img_w = 1, 7, 128, 27, 48, 120, 360, 1080, 1920
    d.mode = _lib.MEAN
    per_segment = 7 * 8 * 16 + (7 * 8 + 8) * 16 + ((7 * 20 + 12 + 15) // 16) * 16      # weights, load list, masks (csrc/ipm_run.cuh)
    assert L.bevipm_plan_bytes(ctypes.byref(d)) == 4096 + 45 * 120 * per_segment
    d.Hb = 121                                                                            # rows are padded to the CTA's four
    assert L.bevipm_plan_bytes(ctypes.byref(d)) == 4096 + 45 * 124 * per_segment
    d.V = 33
    assert L.bevipm_plan_bytes(ctypes.byref(d)) == -1
    d.V, d.Hb = 7, 120
    assert L.bevipm_warp_fuse_fwd_planned(ctypes.byref(d), p16, p16, p16, p16, p16, p16, None, 0, None) == -1  # no cache buffer
    assert L.bevipm_warp_fuse_fwd_planned(ctypes.byref(d), p16, p16, p16, p16, p16, p16, p16, 1024, None) == -1  # too small
    d.variant = 21
    assert L.bevipm_warp_fuse_fwd_planned(ctypes.byref(d), p16, p16, p16, p16, p16, p16, p16, 1 << 30, None) == -1  # default kernels only


def test_kornia_decision_mirrors_the_reference():
    """geometry.py:5-9 + :124: the kornia branch runs iff warp_impl == 'kornia' AND kornia imports.  kornia is not installed
    in this image, so warp_impl='kornia' (what BEVNet passes, model_wrapper.py:42) resolves to the grid_sample geometry."""
    import bevipm
    from bevipm import modules
    try:
        import kornia  # noqa: F401
        have = True
    except Exception:
        have = False
    g = bevipm.GeometryTransformer(8, 8, (-1.0, 1.0, -1.0, 1.0), "kornia")
    assert g.emulate_kornia == have
    assert bevipm.GeometryTransformer(8, 8, (-1.0, 1.0, -1.0, 1.0), "grid_sample", emulate_kornia=True).emulate_kornia is False
    with pytest.warns(UserWarning):
        modules._KORNIA_NOTE = False
        assert bevipm.FusedIPM(8, 8, (-1.0, 1.0, -1.0, 1.0), warp_impl="kornia", emulate_kornia=True).emulate_kornia is True
    # singular feature->BEV matrices are found per view (geometry.py:134-136): a K with a zero focal length
    K = torch.eye(3)[None, None].repeat(1, 2, 1, 1)
    K[0, :, 0, 0] = K[0, :, 1, 1] = 1000.0
    K[0, 1, 0, 0] = 0.0
    Rt = torch.eye(4)[None, None, :3].repeat(1, 2, 1, 1)
    Rt[..., 2, 3] = 5.0
    sing = g._kornia_singular(K, Rt, (20, 30), (1080, 1920))
    assert sing.tolist() == [[False, True]]
