"""GPU: the drop-in voxel_pooling operator vs the oracle and vs the reference's own CUDA kernel."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu

U = 2.0 ** -24  # fp32 unit roundoff


def _sum_tolerance(want64, geom, feat, X, Y, Z):
    """North-star tolerance (rtol 1e-5 / atol 1e-5) plus the fp32 accumulation noise floor of a
    sequential sum, 4 * u * sum|x_i| per voxel: it only matters for stress voxels holding >1000
    points, where the reference's own atomicAdd result is just as far from the fp64 sum."""
    l1, _ = CO.voxel_pooling_forward(geom, np.abs(feat), X, Y, Z, acc64=True)
    return 1e-5 + 1e-5 * np.abs(want64) + 4 * U * l1


def _random_case(B, N, C, X, Y, Z, seed, frac_oob=0.25):
    g = torch.Generator().manual_seed(seed)
    geom = torch.stack([torch.randint(0, X, (B, N), generator=g), torch.randint(0, Y, (B, N), generator=g),
                        torch.randint(0, Z, (B, N), generator=g)], -1).int()
    # heavy-tailed occupancy: a third of the points pile into a few voxels
    hot = torch.rand(B, N, generator=g) < 0.33
    geom[..., 0][hot] = geom[..., 0][hot] % 3
    geom[..., 1][hot] = geom[..., 1][hot] % 2
    oob = torch.rand(B, N, generator=g) < frac_oob
    which = torch.randint(0, 6, (B, N), generator=g)
    vals = torch.tensor([-1, -7, X, Y + 3, Z, -2147483648])
    axis = torch.tensor([0, 1, 0, 1, 2, 0])
    for k in range(6):
        sel = oob & (which == k)
        geom[..., int(axis[k])][sel] = int(vals[k])
    feat = torch.randn(B, N, C, generator=g)
    return geom.contiguous(), feat.contiguous()


CASES = [(2, 5000, 80, 128, 128, 1), (1, 4097, 87, 16, 24, 2), (3, 2048, 5, 8, 8, 1),
         (1, 30000, 128, 64, 32, 1), (2, 3000, 256, 16, 16, 3), (1, 1, 7, 4, 4, 1), (1, 777, 33, 352, 352, 1)]


@pytest.mark.parametrize("B,N,C,X,Y,Z", CASES)
def test_forward_backward_vs_oracle(B, N, C, X, Y, Z):
    from sgv3d_b200 import voxel_pooling
    geom, feat = _random_case(B, N, C, X, Y, Z, seed=N + C)
    f = feat.cuda().requires_grad_(True)
    bev = voxel_pooling(geom.cuda(), f, torch.tensor([X, Y, Z]))
    assert bev.shape == (B, C, Y, X)
    assert bev.permute(0, 2, 3, 1).is_contiguous()          # permuted view, voxel_pooling.py:55
    out_o, pos_o = CO.voxel_pooling_forward(geom.numpy(), feat.numpy(), X, Y, Z)
    got = bev.detach().permute(0, 2, 3, 1).cpu().numpy()
    # ordered per-voxel sums == the oracle's sequential point-order sums: bitwise
    assert np.array_equal(got.view(np.int32), out_o.view(np.int32))
    out64, _ = CO.voxel_pooling_forward(geom.numpy(), feat.numpy(), X, Y, Z, acc64=True)
    assert (np.abs(got - out64) <= _sum_tolerance(out64, geom.numpy(), feat.numpy(), X, Y, Z)).all()
    # backward through .contiguous() (planar gradient) and directly (channels-last gradient)
    gb = torch.randn(B, C, Y, X, generator=torch.Generator().manual_seed(1))
    want = CO.voxel_pooling_backward(gb.numpy(), pos_o, C)
    bev.contiguous().backward(gb.cuda())
    assert np.array_equal(f.grad.cpu().numpy(), want)
    f.grad = None
    bev2 = voxel_pooling(geom.cuda(), f, [X, Y, Z])
    bev2.backward(gb.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    assert np.array_equal(f.grad.cpu().numpy(), want)


@pytest.mark.parametrize("B,N,C,X,Y,Z", CASES[:4])
def test_forward_vs_reference_kernel(B, N, C, X, Y, Z):
    """The reference's own kernel (compiled unmodified into oracle/_ref) on the same inputs:
    same pos_memo bit for bit, features within fp32 tolerance (its atomics reorder the sums)."""
    if not CO.reference_kernel_available():
        pytest.skip("oracle/_ref not built")
    from sgv3d_b200 import _native as Nn
    geom, feat = _random_case(B, N, C, X, Y, Z, seed=7 * N + C)
    gd, fd = geom.cuda(), feat.cuda()
    out_r = torch.zeros(B, Y, X, C, device="cuda")
    pos_r = torch.full((B, N, 3), -1, dtype=torch.int32, device="cuda")
    CO.reference_voxel_pooling_forward(B, N, C, X, Y, Z, gd.data_ptr(), fd.data_ptr(), out_r.data_ptr(),
                                       pos_r.data_ptr(), torch.cuda.current_stream().cuda_stream)
    L = Nn.lib()
    out = torch.empty(B, Y, X, C, device="cuda")
    pos = torch.empty(B, N, 3, dtype=torch.int32, device="cuda")
    wsb = L.sgv3d_voxel_pooling_workspace_bytes(B, N, C, X, Y, Z)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    Nn.check(L.sgv3d_voxel_pooling_forward(B, N, C, X, Y, Z, gd.data_ptr(), fd.data_ptr(), out.data_ptr(),
                                           pos.data_ptr(), ws.data_ptr(), wsb, Nn.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(pos, pos_r)
    out64, _ = CO.voxel_pooling_forward(geom.numpy(), feat.numpy(), X, Y, Z, acc64=True)
    tol = _sum_tolerance(out64, geom.numpy(), feat.numpy(), X, Y, Z)
    assert (np.abs(out.cpu().numpy() - out64) <= tol).all()
    assert (np.abs(out_r.cpu().numpy() - out64) <= tol).all()      # the reference obeys the same bound
    assert (np.abs(out.cpu().numpy() - out_r.cpu().numpy()) <= 2 * tol).all()


def test_deterministic_and_fully_written():
    from sgv3d_b200 import _native as Nn
    B, N, C, X, Y, Z = 2, 20000, 80, 32, 32, 1
    geom, feat = _random_case(B, N, C, X, Y, Z, seed=3)
    gd, fd = geom.cuda(), feat.cuda()
    L = Nn.lib()
    wsb = L.sgv3d_voxel_pooling_workspace_bytes(B, N, C, X, Y, Z)
    outs = []
    for fill in (float("nan"), 7.0, -1.0):
        out = torch.full((B, Y, X, C), fill, device="cuda")
        pos = torch.full((B, N, 3), 12345, dtype=torch.int32, device="cuda")
        ws = torch.randint(0, 255, (wsb,), dtype=torch.uint8, device="cuda")   # dirty workspace
        Nn.check(L.sgv3d_voxel_pooling_forward(B, N, C, X, Y, Z, gd.data_ptr(), fd.data_ptr(), out.data_ptr(),
                                               pos.data_ptr(), ws.data_ptr(), wsb, Nn.current_stream()))
        outs.append((out.clone(), pos.clone()))
    for o, p in outs[1:]:
        assert torch.equal(o.view(torch.int32), outs[0][0].view(torch.int32))
        assert torch.equal(p, outs[0][1])
    assert not torch.isnan(outs[0][0]).any()


def test_all_points_dropped_and_empty():
    from sgv3d_b200 import voxel_pooling
    geom = torch.full((2, 100, 3), -5, dtype=torch.int32, device="cuda")
    feat = torch.randn(2, 100, 16, device="cuda", requires_grad=True)
    bev = voxel_pooling(geom, feat, torch.tensor([8, 4, 1]))
    assert bev.shape == (2, 16, 4, 8) and float(bev.detach().abs().max()) == 0.0
    bev.contiguous().sum().backward()
    assert float(feat.grad.abs().max()) == 0.0
    # multi-dim leading shape like the call site: (B, Nc, D, H, W, 3|C)
    geom6 = torch.randint(0, 4, (1, 2, 3, 2, 5, 3), dtype=torch.int32, device="cuda")
    feat6 = torch.randn(1, 2, 3, 2, 5, 6, device="cuda", requires_grad=True)
    bev6 = voxel_pooling(geom6, feat6, torch.tensor([4, 4, 4]).cuda())
    bev6.contiguous().sum().backward()
    assert feat6.grad.shape == feat6.shape and float(feat6.grad.min()) == 1.0


def test_error_behaviour_matches_reference():
    from sgv3d_b200 import voxel_pooling
    geom = torch.zeros(1, 10, 3, dtype=torch.int32, device="cuda")
    feat = torch.zeros(1, 10, 4, device="cuda")
    with pytest.raises(AssertionError):      # voxel_pooling.py:25-26
        voxel_pooling(geom.expand(2, 10, 3)[:, ::2], feat, [4, 4, 1])
    with pytest.raises(AssertionError):      # voxel_pooling.py:33
        voxel_pooling(geom, torch.zeros(1, 9, 4, device="cuda"), [4, 4, 1])
    with pytest.raises(RuntimeError):        # voxel_pooling_forward.cpp:12-18 CHECK_CUDA
        voxel_pooling(geom.cpu(), feat.cpu(), [4, 4, 1])
    with pytest.raises(RuntimeError):        # data_ptr<int>() dtype mismatch, .cpp:30
        voxel_pooling(geom.long(), feat, [4, 4, 1])
