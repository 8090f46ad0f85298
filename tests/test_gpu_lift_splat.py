"""GPU: fused lift-splat (plan / forward / backward) vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import lift_splat_oracle as O
from sgv3d_b200 import get_shape
from sgv3d_b200.synthetic import make_activations, make_mats
from tests.helpers import frustum_axes, kept_mask_np, oracle_frustum

pytestmark = pytest.mark.gpu


class _BlockWhereSupported(int):
    """pipeline selector resolved per plan: BLOCK when the shape is supported (C <= 96, D <= 255), else TILE"""


@pytest.fixture(autouse=True, params=["tile", "block"])
def pipeline(request, monkeypatch):
    """Every test of this file runs on both kernel pipelines: "tile" = voxel-tile (global sort by voxel + row gathers;
    the default), "block" = the pixel-block pipeline wherever it supports the shape, else voxel-tile."""
    from sgv3d_b200 import view_transform as VT
    if request.param == "tile":
        VT.set_default_pipeline(VT.PIPELINE_TILE)
    else:
        VT.set_default_pipeline(VT.PIPELINE_BLOCK)
        orig = VT.LiftSplatPlan.__init__

        def init(self, frustum, *a, **kw):
            ch = a[9] if len(a) > 9 else kw.get("channels")
            d = int(frustum.shape[0])
            if ch is None or ch > 96 or d > 255:      # not supported by the pixel-block pipeline
                kw["pipeline"] = VT.PIPELINE_TILE
            return orig(self, frustum, *a, **kw)
        monkeypatch.setattr(VT.LiftSplatPlan, "__init__", init)
    yield request.param
    VT.set_default_pipeline(VT.PIPELINE_AUTO)

RTOL, ATOL = 1e-5, 1e-5           # fp32 tolerance stated by BASELINE.json:north_star
RTOL_BF16, ATOL_BF16 = 1e-2, 1e-3  # bf16-context variant vs the fp32 oracle fed the same bf16-rounded inputs


def _setup(shape_name, batch, num_cams, seed, bda, arith=0, ctx_dtype=torch.float32, peaky=False):
    from sgv3d_b200.view_transform import LiftSplatPlan
    shape = get_shape(shape_name) if isinstance(shape_name, str) else shape_name
    mats = make_mats(shape, batch, num_cams, seed=seed, bda=bda)
    fr = oracle_frustum(shape)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    dev = {k: (v.cuda() if v is not None else None) for k, v in mats.items()}
    plan = LiftSplatPlan(fr, dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                         dev["reference_heights"], dev["bda"], vc, vs, shape.grid, shape.channels,
                         ctx_dtype=ctx_dtype, arith=arith)
    # oracle indices from the same 4x4 operands
    ida_inv, mv, me = O.camera_matrices(dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"])
    u, v, z = (t.numpy() for t in frustum_axes(fr))
    bdan = mats["bda"].numpy() if mats["bda"] is not None else None
    xyz = CO.geometry(arith, u, v, z, ida_inv.cpu().numpy(), mv.cpu().numpy(), me.cpu().numpy(),
                      mats["reference_heights"].numpy(), bdan)
    idx = CO.quantize(xyz, (vc - vs / 2.0).numpy(), vs.numpy())
    logits, ctx = make_activations(shape, batch, num_cams, seed=seed, peaky=peaky)
    height = logits.softmax(1)
    return shape, plan, idx, height, ctx


CASES = [("tiny", 2, 2, None), ("small", 3, 1, "random"), ("dair_r50", 2, 1, "identity"),
         ("rope3d_r50", 1, 1, "identity"), ("sgv3d_bsm_r50", 1, 1, "identity"), ("dair_r50_256", 1, 1, "identity"),
         # D = 180 / 256 x 256 and 352 x 352 grids (exps/bevheight/rope3d/bev_height_lss_r101_864_1536_256x256.py:45-54,
         # exps/sgv3d/bsm_bev_height_lss_r101_864_1536_256x256.py:43-46, ...r101_140.8...:45-54), random BDA, 2 cameras
         ("rope3d_r101_256", 1, 1, "identity"), ("sgv3d_bsm_r101", 1, 1, "identity"), ("rope3d_r101_140", 1, 1, "random"),
         ("rope3d_native", 1, 2, "random")]


@pytest.mark.parametrize("shape_name,batch,num_cams,bda", CASES)
def test_plan_expands_to_oracle_voxels(shape_name, batch, num_cams, bda):
    """kept mask + voxel id of every point, recovered from the sorted run index: bit-exact."""
    shape, plan, idx, _, _ = _setup(shape_name, batch, num_cams, 31, bda)
    X, Y, Z = shape.grid
    kept = kept_mask_np(idx, shape.grid)
    want = np.where(kept, idx[..., 1] * X + idx[..., 0], -1).astype(np.int32)
    got = plan.expand().cpu().numpy()
    assert int((got != want).sum()) == 0


@pytest.mark.parametrize("shape_name,batch,num_cams,bda", CASES)
def test_forward_backward_vs_fp64_oracle(shape_name, batch, num_cams, bda):
    shape, plan, idx, height, ctx = _setup(shape_name, batch, num_cams, 32, bda, peaky=(shape_name == "small"))
    X, Y, Z = shape.grid
    bev = plan.forward(height.cuda(), ctx.cuda())
    assert bev.shape == (batch, shape.channels, Y, X) and bev.is_contiguous()
    want = CO.lift_splat_forward64(idx, height.numpy(), ctx.numpy(), X, Y, Z)
    np.testing.assert_allclose(bev.cpu().numpy(), want, rtol=RTOL, atol=ATOL)
    gb = torch.randn(bev.shape, generator=torch.Generator().manual_seed(5))
    g_h, g_c = plan.backward(gb.cuda(), height.cuda(), ctx.cuda())
    gh64, gc64 = CO.lift_splat_backward64(idx, height.numpy(), ctx.numpy(), gb.numpy(), X, Y, Z)
    np.testing.assert_allclose(g_h.cpu().numpy(), gh64, rtol=RTOL, atol=ATOL * 10)   # |sum of C terms| ~ sqrt(C)
    np.testing.assert_allclose(g_c.cpu().numpy(), gc64, rtol=RTOL, atol=ATOL)


def test_matches_reference_port_end_to_end():
    """Whole call site incl. softmax vs the torch-CPU port of _forward_single_sweep (lss_fpn.py:462-495),
    and autograd gradients w.r.t. the raw height-net output."""
    from sgv3d_b200 import LiftSplat
    shape = get_shape("small")
    B = 2
    mats = make_mats(shape, B, 1, seed=40, bda="identity")
    logits, ctx = make_activations(shape, B, 1, seed=40)
    fr = oracle_frustum(shape)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    gb = torch.randn(B, shape.channels, shape.grid[1], shape.grid[0], generator=torch.Generator().manual_seed(2))
    # the oracle runs the 4x4 prep on CPU; the module runs it on the GPU.  Compare indices first.
    bev_o, gl_o, gc_o = O.lift_splat_forward_backward(logits, ctx, fr, mats, vc, vs, vn, gb)
    mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                    shape.downsample, shape.channels).cuda()
    md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).cuda(), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).cuda(),
          "intrin_mats": mats["intrin"].unsqueeze(1).cuda(), "ida_mats": mats["ida"].unsqueeze(1).cuda(),
          "reference_heights": mats["reference_heights"].unsqueeze(1).cuda(), "bda_mat": mats["bda"].cuda()}
    hf = torch.cat((logits, ctx), 1).cuda().requires_grad_(True)
    bev = mod.forward_single_sweep(hf, md)
    bev.backward(gb.cuda())
    idx_o = O.quantize(O.geometry_matmul(fr, mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"], mats["ida"],
                                         mats["reference_heights"], mats["bda"]), vc, vs).numpy()
    idx_k = mod.get_geometry_indices(md["sensor2ego_mats"][:, 0], md["sensor2virtual_mats"][:, 0], md["intrin_mats"][:, 0],
                                     md["ida_mats"][:, 0], md["reference_heights"][:, 0], md["bda_mat"]).cpu().numpy()
    flips = int((idx_k != idx_o).any(-1).sum())
    if flips == 0:   # GPU 4x4 inverses may differ from LAPACK's in the last ulp; only compare when the index sets agree
        torch.testing.assert_close(bev.cpu(), bev_o, rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(hf.grad[:, :shape.D].cpu(), gl_o, rtol=1e-4, atol=ATOL)
        torch.testing.assert_close(hf.grad[:, shape.D:].cpu(), gc_o, rtol=RTOL, atol=ATOL)
    else:
        assert flips < 1e-3 * idx_o.size


def test_bsm_call_site():
    """BSMLSSFPN variant: 80 context + 7 semantic softmax channels, background mask (bsm_lss_fpn.py:523-559)."""
    from sgv3d_b200 import LiftSplat
    shape = get_shape("small")
    B = 2
    mats = make_mats(shape, B, 1, seed=41, bda="identity")
    g = torch.Generator().manual_seed(9)
    height_logits = torch.randn(B, shape.D, shape.fH * 2, shape.fW * 2, generator=g)
    semantic_logits = torch.randn(B, 7, shape.fH * 2, shape.fW * 2, generator=g) * 2
    context = torch.randn(B, 80, shape.fH * 2, shape.fW * 2, generator=g)
    mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                    shape.downsample, 87, is_bsm=True).cuda()
    assert tuple(mod.frustum.shape[1:3]) == (shape.fH * 2, shape.fW * 2)
    md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).cuda(), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).cuda(),
          "intrin_mats": mats["intrin"].unsqueeze(1).cuda(), "ida_mats": mats["ida"].unsqueeze(1).cuda(),
          "reference_heights": mats["reference_heights"].unsqueeze(1).cuda(), "bda_mat": mats["bda"].cuda()}
    bev = mod.forward_single_sweep_bsm(height_logits.cuda(), semantic_logits.cuda(), context.cuda(), md)
    assert bev.shape == (B, 87, shape.grid[1], shape.grid[0])
    # oracle: same indices as the module, context assembled by the port, fp64 accumulation
    idx = mod.get_geometry_indices(md["sensor2ego_mats"][:, 0], md["sensor2virtual_mats"][:, 0], md["intrin_mats"][:, 0],
                                   md["ida_mats"][:, 0], md["reference_heights"][:, 0], md["bda_mat"]).cpu().numpy()
    feat = O.bsm_context(context, semantic_logits)
    want = CO.lift_splat_forward64(idx, height_logits.softmax(1).numpy(), feat.numpy(), *shape.grid)
    np.testing.assert_allclose(bev.cpu().numpy(), want, rtol=RTOL, atol=ATOL)


def test_bf16_context():
    shape, plan, idx, height, ctx = _setup("dair_r50", 1, 1, 33, "identity", ctx_dtype=torch.bfloat16)
    X, Y, Z = shape.grid
    cb = ctx.bfloat16()
    bev = plan.forward(height.cuda(), cb.cuda())
    want = CO.lift_splat_forward64(idx, height.numpy(), cb.float().numpy(), X, Y, Z)
    np.testing.assert_allclose(bev.cpu().numpy(), want, rtol=RTOL_BF16, atol=ATOL_BF16)
    gb = torch.randn(bev.shape, generator=torch.Generator().manual_seed(6))
    g_h, g_c = plan.backward(gb.cuda(), height.cuda(), cb.cuda())
    gh64, gc64 = CO.lift_splat_backward64(idx, height.numpy(), cb.float().numpy(), gb.numpy(), X, Y, Z)
    np.testing.assert_allclose(g_h.cpu().numpy(), gh64, rtol=RTOL_BF16, atol=ATOL_BF16)
    np.testing.assert_allclose(g_c.cpu().numpy(), gc64, rtol=RTOL_BF16, atol=ATOL_BF16)


def test_bitwise_deterministic():
    shape, plan, idx, height, ctx = _setup("rope3d_r50", 2, 1, 34, "identity")
    h, c = height.cuda(), ctx.cuda()
    gb = torch.randn(2, shape.channels, shape.grid[1], shape.grid[0], device="cuda")
    ref = None
    for _ in range(3):
        plan.rebuild()
        out = (plan.forward(h, c), *plan.backward(gb, h, c))
        if ref is None:
            ref = [t.clone() for t in out]
        else:
            for a, b in zip(out, ref):
                assert torch.equal(a.view(torch.int32), b.view(torch.int32))


def test_fused_equals_op_level_path():
    """fused lift-splat == geometry_indices + materialised frustum features + drop-in voxel_pooling."""
    from sgv3d_b200 import geometry_indices, voxel_pooling
    shape, plan, idx, height, ctx = _setup("small", 2, 1, 35, "random")
    bev = plan.forward(height.cuda(), ctx.cuda())
    feat = O.lift(height, ctx).reshape(2, 1, shape.channels, shape.D, shape.fH, shape.fW).permute(0, 1, 3, 4, 5, 2)
    bev_op = voxel_pooling(torch.from_numpy(idx).cuda(), feat.contiguous().cuda(), list(shape.grid)).contiguous()
    torch.testing.assert_close(bev, bev_op, rtol=RTOL, atol=ATOL)


def test_inverse_without_host_sync_is_bit_identical():
    """camera_matrices uses linalg.inv_ex (no info check => no host sync); it must give the very bits
    torch.inverse / Tensor.inverse give, because those 16 floats feed the bit-exact geometry."""
    from sgv3d_b200.view_transform import _inverse
    shape = get_shape("dair_r50")
    m = make_mats(shape, 64, 2, seed=77, bda="random")
    for key in ("sensor2ego", "sensor2virtual", "intrin", "ida"):
        x = m[key].cuda()
        assert torch.equal(_inverse(x), torch.inverse(x)), key
        assert torch.equal(_inverse(x), x.inverse()), key


def test_fused_softmax_on_strided_head_output():
    """logits=True (softmax over D fused, forward and backward) on channel-slice views of a wider
    (BN, D + C + extra, fH, fW) tensor == torch softmax + the probability path on contiguous copies."""
    from sgv3d_b200 import lift_splat
    shape, plan, idx, _, _ = _setup("rope3d_r50", 2, 1, 36, "identity")
    D, C = shape.D, shape.channels
    g = torch.Generator(device="cuda").manual_seed(3)
    big = torch.randn(2, D + C + 5, shape.fH, shape.fW, device="cuda", generator=g)
    big[:, :D] *= 3.0
    a = big.clone().requires_grad_(True)
    bev_a = lift_splat(a[:, :D], a[:, D:D + C], plan, logits=True)
    b = big.clone().requires_grad_(True)
    bev_b = lift_splat(b[:, :D].softmax(1).contiguous(), b[:, D:D + C].contiguous(), plan)
    torch.testing.assert_close(bev_a, bev_b, rtol=1e-5, atol=1e-5)
    gb = torch.randn(bev_a.shape, device="cuda", generator=g)
    bev_a.backward(gb)
    bev_b.backward(gb)
    torch.testing.assert_close(a.grad[:, D:D + C], b.grad[:, D:D + C], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(a.grad[:, :D], b.grad[:, :D], rtol=1e-4, atol=1e-5)
    assert float(a.grad[:, D + C:].abs().max()) == 0.0
    # fp64 anchor for the fused-softmax forward
    want = CO.lift_splat_forward64(idx, big[:, :D].double().softmax(1).float().cpu().numpy(),
                                   big[:, D:D + C].cpu().numpy(), *shape.grid)
    np.testing.assert_allclose(bev_a.detach().cpu().numpy(), want, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("shape_name,batch,arith", [("dair_r50", 48, 2), ("rope3d_r50", 32, 2), ("dair_r50_256", 16, 2),
                                                    ("rope3d_r101_140", 8, 0), ("sgv3d_bsm_r50", 8, 1),
                                                    ("rope3d_native", 16, 2)])
def test_plan_shortcuts_match_the_exact_geometry_kernel(shape_name, batch, arith):
    """Full-size, many calibrations: the plan's fast index paths (z-preserving IDA, identity BDA, guarded
    linear shortcut) against the standalone geometry kernel, which always evaluates the full chain.
    Bit-exact voxel id / kept mask for every point (tens of millions per case)."""
    from sgv3d_b200.view_transform import LiftSplatPlan, geometry_indices
    shape = get_shape(shape_name)
    fr = oracle_frustum(shape)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    X, Y, Z = shape.grid
    total = 0
    for seed, bda in ((101, "identity"), (202, None)):
        mats = make_mats(shape, batch, 1, seed=seed, bda=bda)
        dev = {k: (v.cuda() if v is not None else None) for k, v in mats.items()}
        args = (dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"], dev["reference_heights"], dev["bda"])
        plan = LiftSplatPlan(fr, *args, vc, vs, shape.grid, shape.channels, arith=arith)
        idx = geometry_indices(fr, *args, vc, vs, arith=arith)
        kept = ((idx[..., 0] >= 0) & (idx[..., 0] < X) & (idx[..., 1] >= 0) & (idx[..., 1] < Y)
                & (idx[..., 2] >= 0) & (idx[..., 2] < Z))
        want = torch.where(kept, idx[..., 1] * X + idx[..., 0], torch.full_like(idx[..., 0], -1))
        got = plan.expand()
        assert int((got != want).sum()) == 0
        total += got.numel()
    assert total >= 2 * batch * shape.points_per_frame


@pytest.mark.parametrize("batch", [1, 2, 7, 32])
def test_stacked_inverse_is_bit_identical(batch):
    """camera_matrices() inverts ida / intrin / sensor2virtual with ONE batched call; every 4x4 must keep the
    bits the reference's three separate ``inverse`` calls produce (lss_fpn.py:361,367,392)."""
    from sgv3d_b200.view_transform import camera_matrices
    for seed, bda in ((3, "identity"), (4, "random")):
        m = make_mats(get_shape("rope3d_r50"), batch, 2, seed=seed, bda=bda)
        s2e, s2v, k, ida = (m[n].cuda() for n in ("sensor2ego", "sensor2virtual", "intrin", "ida"))
        got = camera_matrices(s2e, s2v, k, ida)
        want = (ida.inverse(), s2v.matmul(torch.inverse(k)), s2e.matmul(torch.inverse(s2v)))
        for g, w in zip(got, want):
            assert torch.equal(g.contiguous().view(torch.int32), w.contiguous().view(torch.int32))


@pytest.mark.parametrize("channels", [3, 32, 33, 64, 96, 97, 128, 160, 200, 256])
def test_channel_sweep_covers_every_row_layout(channels):
    """BASELINE config 5 sweeps 64-256 channels: every (lanes per row, vectors per lane) instantiation of the
    reduce / backward kernels against the fp64 oracle, with the softmax fused (logits in)."""
    from sgv3d_b200.view_transform import LiftSplatPlan
    shape = get_shape("small")
    B = 2
    mats = make_mats(shape, B, 1, seed=40 + channels, bda="identity")
    fr = oracle_frustum(shape)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    dev = {k: v.cuda() for k, v in mats.items()}
    plan = LiftSplatPlan(fr, dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                         dev["reference_heights"], dev["bda"], vc, vs, shape.grid, channels, arith=0)
    ida_inv, mv, me = O.camera_matrices(dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"])
    u, v, z = (t.numpy() for t in frustum_axes(fr))
    xyz = CO.geometry(0, u, v, z, ida_inv.cpu().numpy(), mv.cpu().numpy(), me.cpu().numpy(),
                      mats["reference_heights"].numpy(), mats["bda"].numpy())
    idx = CO.quantize(xyz, (vc - vs / 2.0).numpy(), vs.numpy())
    logits, ctx = make_activations(shape, B, 1, seed=channels, channels=channels)
    X, Y, Z = shape.grid
    bev = plan.forward(logits.cuda(), ctx.cuda(), logits=True)
    height = logits.softmax(1)
    want = CO.lift_splat_forward64(idx, height.numpy(), ctx.numpy(), X, Y, Z)
    np.testing.assert_allclose(bev.cpu().numpy(), want, rtol=RTOL, atol=ATOL)
    gb = torch.randn(bev.shape, generator=torch.Generator().manual_seed(channels))
    g_h, g_c = plan.backward(gb.cuda(), height.cuda(), ctx.cuda())
    gh64, gc64 = CO.lift_splat_backward64(idx, height.numpy(), ctx.numpy(), gb.numpy(), X, Y, Z)
    np.testing.assert_allclose(g_h.cpu().numpy(), gh64, rtol=RTOL, atol=ATOL * 10)
    np.testing.assert_allclose(g_c.cpu().numpy(), gc64, rtol=RTOL, atol=ATOL)


def test_static_calibration_graph_and_refresh():
    """LiftSplatGraph(static_calibration=True): plan built once outside the graph (static roadside camera),
    replays bit-identical to the per-step path; refresh_calibration() swaps the matrices in place."""
    from sgv3d_b200 import LiftSplat, LiftSplatGraph
    shape = get_shape("small")
    B = 2

    def md_of(seed):
        mats = make_mats(shape, B, 1, seed=seed, bda="identity")
        return {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).cuda(),
                "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).cuda(),
                "intrin_mats": mats["intrin"].unsqueeze(1).cuda(), "ida_mats": mats["ida"].unsqueeze(1).cuda(),
                "reference_heights": mats["reference_heights"].unsqueeze(1).cuda(), "bda_mat": mats["bda"].cuda()}

    mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                    shape.downsample, shape.channels).cuda()
    md_a, md_b = md_of(71), md_of(72)
    logits, ctx = make_activations(shape, B, 1, seed=7)
    hf = torch.cat((logits, ctx), 1).cuda()
    with torch.no_grad():
        want_a = mod.forward_single_sweep(hf, md_a).clone()
        want_b = mod.forward_single_sweep(hf, md_b).clone()
    assert not torch.equal(want_a, want_b)
    g = LiftSplatGraph(mod, hf.clone(), md_a, static_calibration=True)   # the graph owns its input buffer
    assert torch.equal(g(), want_a)
    hf2 = hf * 0.5 + 0.1
    with torch.no_grad():
        want_a2 = mod.forward_single_sweep(hf2, md_a).clone()
    assert torch.equal(g(hf2), want_a2)      # new activations are copied into the captured input
    g.refresh_calibration(md_b)
    assert torch.equal(g(hf), want_b)
    with pytest.raises(RuntimeError):
        g(hf, md_b)     # a different mats_dict must go through refresh_calibration()


@pytest.mark.parametrize("shape_name,batch,bg_shift", [("small", 2, 1.6), ("sgv3d_bsm_r50", 1, 1.6), ("small", 2, 5.0),
                                                       ("sgv3d_bsm_r50", 1, 5.0), ("small", 2, -4.0), ("small", 2, 50.0)])
def test_fused_bsm_assembly_is_bit_identical_to_the_torch_assembly(shape_name, batch, bg_shift):
    """sgv3d_lift_splat_forward_bsm (softmax over the semantic channels + concat + background mask inside the
    context pass, bsm_lss_fpn.py:524-529) == the same torch calls the reference makes, run on the GPU, followed by
    the plain forward: the background mask must be bit-identical, hence the BEV map bitwise equal."""
    from sgv3d_b200 import LiftSplat
    shape = get_shape(shape_name)
    bsm_native = shape_name == "sgv3d_bsm_r50"
    fh, fw = (shape.fH, shape.fW) if bsm_native else (shape.fH * 2, shape.fW * 2)
    mats = make_mats(shape, batch, 1, seed=43, bda="identity")
    g = torch.Generator().manual_seed(11)
    height_logits = torch.randn(batch, shape.D, fh, fw, generator=g).cuda()
    # logits scaled so that the background probability straddles the 0.45 threshold for many pixels
    semantic_logits = (torch.randn(batch, 7, fh, fw, generator=g) * 1.5).cuda()
    semantic_logits[:, 0] += bg_shift   # 1.6: about half of the pixels are background; 5: ~90 % (a roadside frame); 50: all
    context = torch.randn(batch, 80, fh, fw, generator=g).cuda()
    mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                    shape.downsample * (1 if bsm_native else 1), 87, is_bsm=not bsm_native).cuda()
    if bsm_native:   # the named shape already carries the stride-8 map: build the module at that stride
        mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                        shape.downsample, 87).cuda()
    md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).cuda(), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).cuda(),
          "intrin_mats": mats["intrin"].unsqueeze(1).cuda(), "ida_mats": mats["ida"].unsqueeze(1).cuda(),
          "reference_heights": mats["reference_heights"].unsqueeze(1).cuda(), "bda_mat": mats["bda"].cuda()}
    with torch.no_grad():
        fused = mod.forward_single_sweep_bsm(height_logits, semantic_logits, context, md)
    # the reference's own calls (bsm_lss_fpn.py:524-529) on the GPU
    semantic = semantic_logits.softmax(dim=1)
    tran_feat = torch.cat((context, semantic), dim=1)
    mask = semantic[:, 0, :, :].unsqueeze(1) > 0.45
    frac = float(mask.float().mean())
    lo, hi = {1.6: (0.2, 0.8), 5.0: (0.8, 0.99), -4.0: (0.0, 0.05), 50.0: (1.0, 1.0)}[bg_shift]
    assert lo <= frac <= hi, frac          # (1.6: the threshold is exercised on both sides)
    tran_feat = tran_feat * (1 - mask.int())
    plan = mod.make_plan(md, 0, 87)
    want = plan.forward(height_logits, tran_feat.float(), logits=True)
    assert torch.equal(fused, want)
    # and the autograd path (torch assembly) still gives the same values
    hl = height_logits.clone().requires_grad_(True)
    via_autograd = mod.forward_single_sweep_bsm(hl, semantic_logits, context, md)
    assert torch.equal(via_autograd.detach(), want)


def test_inverse4x4_kernel_is_bit_identical_to_torch():
    """sgv3d_inverse4x4 restates the arithmetic of torch.inverse on CUDA (the call the reference makes at
    lss_fpn.py:361,367,392): every bit of every inverse must agree, for calibration matrices of both families,
    random BDA, matrices that need row pivoting and badly scaled ones."""
    from sgv3d_b200.view_transform import _inverse_kernel_verified, camera_matrices, inverse4x4
    dev = torch.device("cuda", 0)
    assert _inverse_kernel_verified(dev)
    sets = []
    for fam in ("dair_r50", "rope3d_r50"):
        m = make_mats(get_shape(fam), 200, 1, seed=61, bda="random")
        sets += [m["ida"], m["intrin"], m["sensor2virtual"], m["sensor2ego"], m["bda"].unsqueeze(1)]
    g = torch.Generator().manual_seed(3)
    rnd = torch.randn(4096, 1, 4, 4, generator=g)
    sets += [rnd, rnd * torch.logspace(-3, 3, 4).view(1, 1, 1, 4), rnd[:, :, [2, 0, 3, 1]] + torch.eye(4)]
    for a in sets:
        a = a.float().cuda()
        want = torch.inverse(a)
        (got,) = inverse4x4(a)
        assert torch.equal(want.view(torch.int32), got.view(torch.int32))
    # three sets in one launch, as camera_matrices uses it, against the reference's three separate calls
    m = make_mats(get_shape("dair_r50"), 37, 2, seed=62, bda=None)
    ida, k, s2v, s2e = (m[n].cuda() for n in ("ida", "intrin", "sensor2virtual", "sensor2ego"))
    ida_inv, mv, me = camera_matrices(s2e, s2v, k, ida)
    assert torch.equal(ida_inv, ida.inverse())
    assert torch.equal(mv, s2v.matmul(torch.inverse(k)))
    assert torch.equal(me, s2e.matmul(torch.inverse(s2v)))


@pytest.mark.parametrize("batch,num_cams", [(1, 1), (2, 1), (1, 3), (7, 1), (32, 1), (64, 2)])
def test_camera_prep_kernel_is_bit_identical_to_torch(batch, num_cams):
    """sgv3d_camera_prep (three inverses + two products, one launch) == the reference's torch calls
    (lss_fpn.py:361,367,392) bit for bit, for every batch count (torch's matmul changes its rounding order between a
    single matrix and a batch)."""
    from sgv3d_b200.view_transform import _camera_prep, _camera_prep_verified, camera_matrices
    m = make_mats(get_shape("rope3d_r50"), batch, num_cams, seed=70 + batch, bda=None)
    ida, k, s2v, s2e = (m[n].cuda() for n in ("ida", "intrin", "sensor2virtual", "sensor2ego"))
    assert _camera_prep_verified(ida.device, ida.shape)
    want = (ida.inverse(), s2v.matmul(torch.inverse(k)), s2e.matmul(torch.inverse(s2v)))
    for got in (_camera_prep(s2e, s2v, k, ida), camera_matrices(s2e, s2v, k, ida)):
        for w, g in zip(want, got):
            assert torch.equal(w.view(torch.int32), g.contiguous().view(torch.int32))


# ---- fused-path edge cases: through plan -> forward -> backward, against the oracle ----------------
def _plan_from(shape, mats, channels=None, ctx_dtype=torch.float32, arith=2):
    from sgv3d_b200.view_transform import LiftSplatPlan
    fr = oracle_frustum(shape)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    dev = {k: (v.cuda() if v is not None else None) for k, v in mats.items()}
    plan = LiftSplatPlan(fr, dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                         dev["reference_heights"], dev["bda"], vc, vs, shape.grid, channels or shape.channels,
                         ctx_dtype=ctx_dtype, arith=arith)
    return plan, fr, vs, vc, dev


def _indices_like_the_gpu(shape, fr, vs, vc, dev, arith=2):
    """voxel indices of the standalone geometry kernel (always the full chain; itself pinned against the C oracle
    and the reference port by tests/test_gpu_geometry.py), as numpy"""
    from sgv3d_b200.view_transform import geometry_indices
    return geometry_indices(fr, dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                            dev["reference_heights"], dev["bda"], vc, vs, arith=arith).cpu().numpy()


def test_fused_path_nan_and_inf_rays():
    """A camera whose virtual-camera ray has pv.y == 0 for a whole image row (height / 0 = +-Inf, 0 / 0 = NaN;
    lss_fpn.py:363) and a BDA that spreads the NaN (lss_fpn.py:394-398): on the GPU `.int()` maps NaN to 0, so such
    points are KEPT in voxel (0, 0, 0).  Plan, forward and backward must agree with the oracle fed the same indices."""
    shape = get_shape("small")
    B = 2
    mats = make_mats(shape, B, 1, seed=51, bda="random")
    # frame 0: sensor2virtual @ K^-1 with a zero second row => pv.y == 0 everywhere => ratio = +-Inf / NaN
    mats["sensor2virtual"][0, 0, 1, :] = 0.0
    # frame 1: reference height NaN => every point NaN
    mats["reference_heights"][1, 0] = float("nan")
    plan, fr, vs, vc, dev = _plan_from(shape, mats)
    idx = _indices_like_the_gpu(shape, fr, vs, vc, dev)
    X, Y, Z = shape.grid
    kept = kept_mask_np(idx, shape.grid)
    want = np.where(kept, idx[..., 1] * X + idx[..., 0], -1).astype(np.int32)
    got = plan.expand().cpu().numpy()
    assert int((got != want).sum()) == 0
    assert kept[1].all() and (want[1] == 0).all()          # NaN rays: kept, voxel 0
    logits, ctx = make_activations(shape, B, 1, seed=51)
    height = logits.softmax(1)
    bev = plan.forward(height.cuda(), ctx.cuda())
    want_bev = CO.lift_splat_forward64(idx, height.numpy(), ctx.numpy(), X, Y, Z)
    np.testing.assert_allclose(bev.cpu().numpy(), want_bev, rtol=RTOL, atol=ATOL)
    gb = torch.randn(bev.shape, generator=torch.Generator().manual_seed(8))
    g_h, g_c = plan.backward(gb.cuda(), height.cuda(), ctx.cuda())
    gh64, gc64 = CO.lift_splat_backward64(idx, height.numpy(), ctx.numpy(), gb.numpy(), X, Y, Z)
    np.testing.assert_allclose(g_h.cpu().numpy(), gh64, rtol=RTOL, atol=ATOL * 10)
    np.testing.assert_allclose(g_c.cpu().numpy(), gc64, rtol=RTOL, atol=ATOL)


def test_fused_path_frame_with_every_point_dropped_and_empty_batch():
    """Frame 1 stands 1 km beyond the grid: no point is kept, its BEV map must be all zeros and its gradients zero;
    frame 0 is a normal frame.  Then B = 0: empty outputs, no launch failure."""
    shape = get_shape("small")
    B = 2
    mats = make_mats(shape, B, 1, seed=52, bda="identity")
    mats["sensor2ego"][1, 0, 0, 3] += 1000.0
    plan, fr, vs, vc, dev = _plan_from(shape, mats)
    idx = _indices_like_the_gpu(shape, fr, vs, vc, dev)
    X, Y, Z = shape.grid
    kept = kept_mask_np(idx, shape.grid)
    assert kept[0].any() and not kept[1].any()
    got = plan.expand().cpu().numpy()
    assert (got[1] == -1).all()
    logits, ctx = make_activations(shape, B, 1, seed=52)
    height = logits.softmax(1)
    bev = plan.forward(height.cuda(), ctx.cuda())
    want_bev = CO.lift_splat_forward64(idx, height.numpy(), ctx.numpy(), X, Y, Z)
    np.testing.assert_allclose(bev.cpu().numpy(), want_bev, rtol=RTOL, atol=ATOL)
    assert float(bev[1].abs().max()) == 0.0
    gb = torch.randn(bev.shape, generator=torch.Generator().manual_seed(9))
    g_h, g_c = plan.backward(gb.cuda(), height.cuda(), ctx.cuda())
    assert float(g_h[1].abs().max()) == 0.0 and float(g_c[1].abs().max()) == 0.0
    gh64, gc64 = CO.lift_splat_backward64(idx, height.numpy(), ctx.numpy(), gb.numpy(), X, Y, Z)
    np.testing.assert_allclose(g_h.cpu().numpy(), gh64, rtol=RTOL, atol=ATOL * 10)
    np.testing.assert_allclose(g_c.cpu().numpy(), gc64, rtol=RTOL, atol=ATOL)
    # B = 0
    empty = {k: (v[:0] if v is not None else None) for k, v in mats.items()}
    plan0, *_ = _plan_from(shape, empty)
    h0 = torch.zeros(0, shape.D, shape.fH, shape.fW, device="cuda")
    c0 = torch.zeros(0, shape.channels, shape.fH, shape.fW, device="cuda")
    bev0 = plan0.forward(h0, c0)
    assert tuple(bev0.shape) == (0, shape.channels, Y, X)
    gh0, gc0 = plan0.backward(torch.zeros(0, shape.channels, Y, X, device="cuda"), h0, c0)
    assert gh0.numel() == 0 and gc0.numel() == 0
    assert plan0.expand().numel() == 0


def test_matches_reference_port_run_on_the_gpu():
    """The torch port of _forward_single_sweep (lss_fpn.py:462-495) executed ON THE GPU -- the device the reference
    runs on -- against the module: voxel indices bit-exact (0 flips), BEV map and gradients within tolerance."""
    from sgv3d_b200 import LiftSplat
    shape = get_shape("small")
    B = 3
    mats = make_mats(shape, B, 1, seed=44, bda="random")
    logits, ctx = make_activations(shape, B, 1, seed=44)
    fr = oracle_frustum(shape)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    gb = torch.randn(B, shape.channels, shape.grid[1], shape.grid[0], generator=torch.Generator().manual_seed(3))
    cm = {k: (v.cuda() if v is not None else None) for k, v in mats.items()}
    bev_o, gl_o, gc_o = O.lift_splat_forward_backward(logits.cuda(), ctx.cuda(), fr.cuda(), cm, vc.cuda(), vs.cuda(),
                                                      vn, gb.cuda())
    idx_o = O.quantize(O.geometry_matmul(fr.cuda(), cm["sensor2ego"], cm["sensor2virtual"], cm["intrin"], cm["ida"],
                                         cm["reference_heights"], cm["bda"]), vc.cuda(), vs.cuda())
    mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                    shape.downsample, shape.channels).cuda()
    md = {"sensor2ego_mats": cm["sensor2ego"].unsqueeze(1), "sensor2virtual_mats": cm["sensor2virtual"].unsqueeze(1),
          "intrin_mats": cm["intrin"].unsqueeze(1), "ida_mats": cm["ida"].unsqueeze(1),
          "reference_heights": cm["reference_heights"].unsqueeze(1), "bda_mat": cm["bda"]}
    idx_k = mod.get_geometry_indices(md["sensor2ego_mats"][:, 0], md["sensor2virtual_mats"][:, 0], md["intrin_mats"][:, 0],
                                     md["ida_mats"][:, 0], md["reference_heights"][:, 0], md["bda_mat"])
    assert int((idx_k != idx_o).sum()) == 0
    hf = torch.cat((logits, ctx), 1).cuda().requires_grad_(True)
    bev = mod.forward_single_sweep(hf, md)
    bev.backward(gb.cuda())
    torch.testing.assert_close(bev, bev_o, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(hf.grad[:, shape.D:], gc_o, rtol=RTOL, atol=ATOL)
    # d/d logits: a sum of D products of O(sqrt(C)) terms through the softmax Jacobian -- rtol 1e-4 (stated in DESIGN.md 2)
    torch.testing.assert_close(hf.grad[:, :shape.D], gl_o, rtol=1e-4, atol=ATOL)


def test_bsm_gradients_vs_oracle():
    """BSMLSSFPN call site under autograd (bsm_lss_fpn.py:523-541): gradients w.r.t. the height logits, the semantic
    logits and the context against torch autograd over the port (O.bsm_context + O.lift + index_add_), fp64."""
    from sgv3d_b200 import LiftSplat
    shape = get_shape("small")
    B = 2
    mats = make_mats(shape, B, 1, seed=45, bda="identity")
    g = torch.Generator().manual_seed(12)
    fh, fw = shape.fH * 2, shape.fW * 2
    hl = torch.randn(B, shape.D, fh, fw, generator=g)
    sl = torch.randn(B, 7, fh, fw, generator=g) * 1.5
    sl[:, 0] += 1.0
    cx = torch.randn(B, 80, fh, fw, generator=g)
    mod = LiftSplat(shape.x_bound, shape.y_bound, shape.z_bound, shape.d_bound, shape.final_dim,
                    shape.downsample, 87, is_bsm=True).cuda()
    md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).cuda(), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).cuda(),
          "intrin_mats": mats["intrin"].unsqueeze(1).cuda(), "ida_mats": mats["ida"].unsqueeze(1).cuda(),
          "reference_heights": mats["reference_heights"].unsqueeze(1).cuda(), "bda_mat": mats["bda"].cuda()}
    a = [t.clone().cuda().requires_grad_(True) for t in (hl, sl, cx)]
    bev = mod.forward_single_sweep_bsm(a[0], a[1], a[2], md)
    gb = torch.randn(bev.shape, generator=torch.Generator().manual_seed(13))
    bev.backward(gb.cuda())
    # oracle: same indices (module's geometry kernel), fp64 autograd over the port
    idx = mod.get_geometry_indices(md["sensor2ego_mats"][:, 0], md["sensor2virtual_mats"][:, 0], md["intrin_mats"][:, 0],
                                   md["ida_mats"][:, 0], md["reference_heights"][:, 0], md["bda_mat"]).cpu()
    o = [t.clone().double().requires_grad_(True) for t in (hl, sl, cx)]
    feat = O.bsm_context(o[2], o[1])
    # the mask is decided in fp32 by the reference; reuse the fp32 decision so that fp64 does not flip a pixel
    mask32 = (sl.softmax(1)[:, 0:1] > 0.45)
    feat = torch.cat((o[2], o[1].softmax(1)), 1) * (1 - mask32.int())
    lifted = O.lift(o[0].softmax(1), feat).reshape(B, 1, 87, shape.D, fh, fw).permute(0, 1, 3, 4, 5, 2)
    bev_o, _ = O.voxel_pooling_forward(idx, lifted.contiguous(), list(shape.grid))
    bev_o = bev_o.contiguous()
    bev_o.backward(gb.double())
    torch.testing.assert_close(bev.detach().cpu().double(), bev_o.detach(), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(a[2].grad.cpu().double(), o[2].grad, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(a[1].grad.cpu().double(), o[1].grad, rtol=1e-4, atol=ATOL)
    torch.testing.assert_close(a[0].grad.cpu().double(), o[0].grad, rtol=1e-4, atol=ATOL)


@pytest.mark.parametrize("seed", range(40))
def test_random_configurations_match_the_oracle(seed):
    """Seeded random problem shapes -- image size, stride, number and range of the height bins, channels, grid extent and
    cell size, z range, cameras per frame, BDA none / identity / random, all three arithmetic orders -- through plan ->
    expand (bit-exact voxel ids), forward and backward (fp64 oracle), on whichever pipeline the fixture selects."""
    from sgv3d_b200.shapes import LiftSplatShape
    rng = np.random.default_rng(1000 + seed)
    stride = int(rng.choice([8, 16]))
    fh, fw = int(rng.integers(2, 12)), int(rng.integers(3, 20))
    cell = float(rng.choice([0.8, 1.6, 3.2]))
    nx, ny = int(rng.integers(6, 70)), int(rng.integers(6, 70))
    z_lo = float(rng.choice([-5.0, -3.0, -1.0]))
    z_hi = float(rng.choice([0.5, 3.0]))
    d_lo = float(rng.choice([-2.0, -1.0, 0.0]))
    shape = LiftSplatShape(f"random{seed}", (fh * stride + int(rng.integers(0, stride)), fw * stride + int(rng.integers(0, stride))),
                           stride, (d_lo, d_lo + float(rng.choice([1.0, 2.0, 5.5])), int(rng.integers(1, 48))),
                           int(rng.integers(1, 41)), x_bound=(0.0, nx * cell, cell), y_bound=(-ny * cell / 2, ny * cell / 2, cell),
                           z_bound=(z_lo, z_hi, z_hi - z_lo), family=str(rng.choice(["dair", "rope3d"])), source="test-only")
    batch, cams = int(rng.integers(1, 4)), int(rng.integers(1, 3))
    bda = [None, "identity", "random"][int(rng.integers(0, 3))]
    arith = int(rng.integers(0, 3))
    shape, plan, idx, height, ctx = _setup(shape, batch, cams, 500 + seed, bda, arith=arith, peaky=bool(seed % 2))
    X, Y, Z = shape.grid
    kept = kept_mask_np(idx, shape.grid)
    want_vox = np.where(kept, idx[..., 1] * X + idx[..., 0], -1).astype(np.int32)
    assert int((plan.expand().cpu().numpy() != want_vox).sum()) == 0, (shape, batch, cams, bda, arith)
    bev = plan.forward(height.cuda(), ctx.cuda())
    want = CO.lift_splat_forward64(idx, height.numpy(), ctx.numpy(), X, Y, Z)
    np.testing.assert_allclose(bev.cpu().numpy(), want, rtol=RTOL, atol=ATOL)
    gb = torch.randn(bev.shape, generator=torch.Generator().manual_seed(seed))
    g_h, g_c = plan.backward(gb.cuda(), height.cuda(), ctx.cuda())
    gh64, gc64 = CO.lift_splat_backward64(idx, height.numpy(), ctx.numpy(), gb.numpy(), X, Y, Z)
    np.testing.assert_allclose(g_h.cpu().numpy(), gh64, rtol=RTOL, atol=ATOL * 10)
    np.testing.assert_allclose(g_c.cpu().numpy(), gc64, rtol=RTOL, atol=ATOL)
