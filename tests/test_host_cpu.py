"""CPU tests of the host side: C-ABI surface, reference-API mirror (names, defaults, errors),
batching arithmetic, and the N>1 partition + psum semantics over gloo (world_size 2)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

from jax_dips_b200 import data_management, mesh, optimizers, problems
from jax_dips_b200 import numpy as jnp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    from jax_dips_b200 import build as b
    lib_path = b.build()                       # nvcc cross-compiles for sm_100a without a GPU
    header = open(os.path.join(ROOT, "include", "nbm_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(nbm_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 15
    L = ctypes.CDLL(lib_path)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    from jax_dips_b200 import _cabi
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    _cabi.lib()
    assert _cabi.lib().nbm_version() >= 100
    # struct layouts agree with the compiler's (sizes as nvcc's host compiler sees them)
    net = _cabi.Net(2, 10, 1, 1)
    assert _cabi.lib().nbm_net_num_params(ctypes.byref(net)) == 167
    net3 = _cabi.Net(2, 10, 1, 3)
    assert _cabi.lib().nbm_net_num_params(ctypes.byref(net3)) == 177          # README.md:137-142


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jax_dips_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle's", ""), fn


def test_mesh_is_z_fastest_like_the_reference():
    init_mesh_fn, coord_at = mesh.construct(3)
    x = torch.linspace(0, 1, 3); y = torch.linspace(0, 2, 4); z = torch.linspace(0, 3, 5)
    g = init_mesh_fn(x, y, z)
    assert g.shape() == (3, 4, 5)
    R = g.R.reshape(3, 4, 5, 3)
    assert torch.equal(R[1, 2, 3], torch.stack((x[1], y[2], z[3])))            # index = (i*Ny + j)*Nz + k
    assert g.R_zmax_boundary.shape == (12, 3) and float(g.R_zmax_boundary[:, 2].min()) == 3.0
    assert float(g.dx) == pytest.approx(0.5)
    assert coord_at(g, (2, 3, 4)) == [x[2], y[3], z[4]]
    with pytest.raises(NotImplementedError):
        mesh.construct(2)


def test_vmap_of_per_point_callables():
    g = mesh.linspace_grid([-1] * 3, [1] * 3, [4, 5, 6])
    f = jnp.vmap(lambda r: jnp.exp(r[2]) * r[1] + jnp.where(r[0] > 0, 1.0, 0.0))
    want = torch.exp(g.R[:, 2]) * g.R[:, 1] + (g.R[:, 0] > 0).float()
    assert torch.allclose(f(g.R), want)
    assert jnp.vmap(lambda r: 0.0)(g.R).shape == (120,)
    P = problems.sphere()
    beta = jnp.vmap(P.beta_fn)(g.R)
    assert beta.shape == (120,) and torch.isfinite(beta).all()


def test_dataset_dict_ranges_follow_the_reference_arithmetic():
    # single device: contiguous batches (data_management.py:121-130)
    DD = data_management.DatasetDict(num_points=64 ** 3, batch_size=131072)
    assert DD.num_batches == 2 and DD.ranges(0) == [(0, 131072), (131072, 262144)]
    # batch larger than the data is clipped (:90-91)
    assert data_management.DatasetDict(num_points=4096, batch_size=131072).ranges(0) == [(0, 4096)]
    # multi device: per-device batch = min(n_dev*batch, ceil(N/n_dev)) (trainer.py:733-737), x-slabs
    DD = data_management.DatasetDict(num_points=32 ** 3, batch_size=8 * 131072, num_gpus=8)
    assert [DD.ranges(r) for r in range(8)] == [[(r * 4096, (r + 1) * 4096)] for r in range(8)]
    # ragged sizes: the short last batch keeps its real points (the reference pads it with PRNGKey(0) random points)
    DD = data_management.DatasetDict(num_points=1000, batch_size=300)
    assert DD.padded and DD.num_batches == 4
    assert DD.ranges(0) == [(0, 300), (300, 600), (600, 900), (900, 1000)]
    # the reference's own LPBE example (lpbe.yaml:76-84): 32^3 points in batches of 3996
    DD = data_management.DatasetDict(num_points=32 ** 3, batch_size=3996)
    r = DD.ranges(0)
    assert len(r) == 9 and r[-1] == (8 * 3996, 32 ** 3) and all(b - a == 3996 for a, b in r[:-1])
    # devices whose block ends early get empty ranges, every device takes the same number of steps
    DD = data_management.DatasetDict(num_points=1001, batch_size=2 * 250, num_gpus=2)
    assert DD.batch_size == 500 and DD.ranges(0) == [(0, 500), (500, 501)] and DD.ranges(1) == [(501, 1001), (1001, 1001)]
    assert not data_management.DatasetDict(num_points=4096, batch_size=1024).padded


def test_zoom_schedule():
    g = mesh.linspace_grid([-1] * 3, [1] * 3, [9, 9, 9])
    TD = data_management.TrainData(g)
    assert [TD.zoom_level(8, e) for e in range(8)] == [0, 0, 1, 1, 2, 2, 3, 3]   # data_management.py:320-326
    assert TD.zoom_level(10, 9) == 4                                             # num_epochs % 4 != 0 reaches 4
    assert TD.zoom_cell(2)[0] == pytest.approx(0.25 * 0.25)
    with pytest.raises(ZeroDivisionError):
        TD.zoom_level(3, 0)


def test_optimizer_factory_mirrors_the_reference():
    o = optimizers.get_optimizer("custom", "exponential", 1e-3, 0.975)
    assert o.kind == 0 and o.scheduler(1000) == pytest.approx(1e-3 * 0.975)
    assert optimizers.get_optimizer("adam", learning_rate=1e-2).kind == 1
    assert optimizers.get_optimizer("rmsprop", learning_rate=1e-2).kind == 2
    with pytest.raises(ValueError):
        optimizers.get_optimizer("sgd")                                           # optimizers.py:95-97
    sig = inspect.signature(optimizers.get_optimizer)
    assert list(sig.parameters)[:5] == ["optimizer_name", "scheduler_name", "learning_rate", "decay_rate", "max_norm"]


def test_trainer_api_signature_matches_the_reference():
    from jax_dips_b200 import trainer
    sig = inspect.signature(trainer.setup)
    assert list(sig.parameters) == ["initial_value_fn", "dirichlet_bc_fn", "lvl_set_fn", "mu_m_fn_", "mu_p_fn_",
                                    "k_m_fn_", "k_p_fn_", "f_m_fn_", "f_p_fn_", "alpha_fn_", "beta_fn_",
                                    "nonlinear_op_m", "nonlinear_op_p"]                 # trainer.py:980-994
    init_fn = trainer.setup(*problems.sphere().setup_args())
    p = inspect.signature(init_fn).parameters
    for name, default in (("num_epochs", 1000), ("batch_size", 131072), ("algorithm", 0), ("multi_gpu", False),
                          ("checkpoint_interval", 1000), ("restart", False), ("print_rate", 1)):
        assert p[name].default == default                                                 # trainer.py:1035-1076
    net = trainer.NetShape.from_model_dict(trainer._DEFAULT_MODEL)
    p0 = trainer.haiku_init(net)
    assert p0.numel() == 167
    tree = trainer.params_to_tree(net, p0)
    assert set(tree) == {"double_mlp/~mlp_p_fn/linear", "double_mlp/~mlp_p_fn/linear_1", "double_mlp/~mlp_p_fn/linear_2",
                         "double_mlp/~mlp_m_fn/linear", "double_mlp/~mlp_m_fn/linear_1", "preconditioner"}
    assert tree["double_mlp/~mlp_p_fn/linear_1"]["w"].shape == (10, 10)
    assert torch.equal(trainer.tree_to_params(net, tree), p0)
    with pytest.raises(NotImplementedError):
        trainer.NetShape.from_model_dict({"model_type": "resnet", "mlp": {}})
    from jax_dips_b200.plan import Nonlinear
    with pytest.raises(NotImplementedError):
        Nonlinear.coerce(lambda u: u ** 3)
    assert Nonlinear.coerce(lambda u: 0.0).kind == 0


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from oracle import nbm_oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    P = problems.sphere()
    tr, lv, phi_grid, op = util.make_case(P, 8, 16, "trilinear", torch.float64)
    DD = data_management.DatasetDict(num_points=tr.num_points(), batch_size=world * 131072, num_gpus=world)
    (p0, p1), = DD.ranges(rank)
    params = O.init_params(op.shape, seed=2, dtype=torch.float64)
    d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    loss, grad = O.loss_and_grad(params, tr.R.double()[p0:p1], *d, op)      # per-device MEAN over its slab
    buf = torch.cat((grad, loss.reshape(1)))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)                               # psum (trainer.py:829-830)
    if rank == 0:
        q.put((p0, p1, buf.numpy()))
    else:
        q.put((p0, p1, None))
    dist.destroy_process_group()


def test_two_rank_partition_and_psum_over_gloo():
    import socket
    import torch.multiprocessing as mp
    import util
    from oracle import nbm_oracle as O
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    ranges = sorted((a, b) for a, b, _ in got)
    assert ranges == [(0, 256), (256, 512)]                                   # contiguous x-slabs of the 8^3 grid
    buf = next(v for _, _, v in got if v is not None)
    # the oracle's own multi-device loop (sum over devices of per-device means)
    P = problems.sphere()
    tr, lv, phi_grid, op = util.make_case(P, 8, 16, "trilinear", torch.float64)
    params = O.init_params(op.shape, seed=2, dtype=torch.float64)
    d = [tr.dx.double(), tr.dy.double(), tr.dz.double()]
    gsum, lsum = torch.zeros_like(params), 0.0
    for dev in range(2):
        l, g = O.loss_and_grad(params, tr.R.double()[dev * 256:(dev + 1) * 256], *d, op)
        gsum += g; lsum += float(l)
    assert np.allclose(buf[:-1], gsum.numpy(), rtol=1e-12, atol=1e-15)
    assert buf[-1] == pytest.approx(lsum, rel=1e-12)
    # and it is NOT the global mean: psum of means = world_size x the single-device mean gradient
    l1, g1 = O.loss_and_grad(params, tr.R.double(), *d, op)
    assert np.allclose(buf[:-1], 2 * g1.numpy(), rtol=1e-9, atol=1e-14)


def test_vtk_writer_roundtrip(tmp_path):
    """write_vtk_manual (jax_dips/utils/io.py:80-89): VTK XML StructuredGrid with appended raw data, x fastest,
    fields given flat in the grid's z-fastest order."""
    from jax_dips_b200 import io as nio
    g = mesh.linspace_grid((-1.0, 0.0, 2.0), (1.0, 3.0, 4.0), [4, 5, 6])
    R = g.R
    sol = (R[:, 0] + 10.0 * R[:, 1] + 100.0 * R[:, 2]).float()
    phi = (R ** 2).sum(1).double()
    path = nio.write_vtk_manual(g, {"sol": sol, "phi": phi}, filename=str(tmp_path / "out" / "dump"))
    assert path.endswith("dump.vts") and os.path.exists(path)
    pts, fields = nio.read_vts(path)
    assert pts.shape == (4 * 5 * 6, 3) and set(fields) == {"sol", "phi"}
    # VTK point order is x fastest; point data must follow it
    P = pts.reshape(6, 5, 4, 3)                     # (k, j, i, xyz)
    assert np.allclose(P[0, 0, :, 0], g.x.numpy()) and np.allclose(P[0, :, 0, 1], g.y.numpy())
    want = (pts[:, 0] + 10.0 * pts[:, 1] + 100.0 * pts[:, 2]).reshape(6, 5, 4).transpose(2, 1, 0)
    assert np.allclose(fields["sol"], want, rtol=1e-6)
    assert fields["phi"].dtype == np.float64
    head = open(path, "rb").read(200).decode("ascii", "replace")
    assert 'type="StructuredGrid"' in head and 'header_type="UInt64"' in head


def test_empty_plan_contributes_zeros():
    """ragged multi-device partitions: a device without points for a batch still joins the exchange with zeros"""
    from jax_dips_b200.plan import EmptyPlan
    pl = EmptyPlan(5, "cpu")
    pl.loss_grad.fill_(3.0)
    out = pl.loss_grad_launch()
    assert out.shape == (6,) and float(out.abs().sum()) == 0.0 and pl.n_points == 0

    class FakeComm:
        def reduce_allreduce(self, partials, rows, np1, target):
            assert rows == 1 and np1 == 6 and float(partials.abs().sum()) == 0.0
            target.copy_(partials + 1.0)      # what the peers contributed
    assert float(pl.loss_grad_launch(comm=FakeComm()).sum()) == 6.0


def test_parameter_tree_roundtrip_with_the_preconditioner():
    """flat vector [network | preconditioner] <-> the checkpoint tree: haiku keys for the network (trainer.py:323), the flax
    tree of nn/preconditioner.py under "preconditioner" (trainer.py:236-243)"""
    from jax_dips_b200 import trainer as ntrainer
    from jax_dips_b200.plan import NetShape, PrecondShape
    net = NetShape()
    pc = PrecondShape.from_model_dict({"preconditioner": {"enable": True, "layer_widths": [8, 4], "scaling_coeff": 2.0}})
    assert pc.scale == 2.0 and pc.n_params == 26 * 8 + 8 + 8 * 4 + 4 + 4 + 1
    assert PrecondShape.from_model_dict({"preconditioner": {"enable": False}}) is None
    assert PrecondShape.from_model_dict({}) is None
    with pytest.raises(NotImplementedError):
        PrecondShape([8, 4, 2])
    flat = torch.cat((ntrainer.haiku_init(net, 1), ntrainer.precond_init(pc, 1)))
    assert flat.numel() == net.n_params + pc.n_params
    tree = ntrainer.params_to_tree(net, flat, pc)
    dense = tree["preconditioner"]["params"]
    assert dense["Dense_0"]["kernel"].shape == (26, 8) and dense["Dense_1"]["kernel"].shape == (8, 4)
    assert dense["Dense_2"]["kernel"].shape == (4, 1) and dense["Dense_2"]["bias"].shape == (1,)
    lim = (6.0 / (26 + 8)) ** 0.5                          # glorot_uniform (nn/preconditioner.py:20)
    assert float(np.abs(dense["Dense_0"]["kernel"]).max()) <= lim and float(np.abs(dense["Dense_0"]["bias"]).max()) == 0.0
    assert torch.equal(ntrainer.tree_to_params(net, tree, pc), flat)
    assert ntrainer.params_to_tree(net, flat[:net.n_params])["preconditioner"] == {}


def test_cabi_argument_validation_without_a_gpu():
    """error convention of the C ABI (include/nbm_b200.h): int status, nbm_last_error(); argument checks come before
    any CUDA call, so they can be exercised on a host without a GPU"""
    import ctypes as C
    from jax_dips_b200 import _cabi as cabi
    L = cabi.lib()
    assert L.nbm_version() >= 100
    net = cabi.Net(2, 10, 1, 1)
    assert L.nbm_net_num_params(C.byref(net)) == 167                   # p 3-10-10-1: 161, m 3-1-1: 6
    assert L.nbm_net_num_params(C.byref(cabi.Net(2, 10, 1, 3))) == 177  # README.md:137-142 of the reference
    assert L.nbm_precond_num_params(8, 4) == 257
    assert L.nbm_step_partial_rows() >= 148
    assert L.nbm_loss_grad_shared_f32(None, None) == 1                  # NBM_ERR_BAD_ARG
    assert b"null" in L.nbm_last_error()
    step = cabi.SharedStep()
    assert L.nbm_loss_grad_shared_f32(C.byref(step), None) == 1
    assert L.nbm_loss_grad_points_f32(None, None) == 1
    opt = cabi.Optimizer(167, 1e-3, 0.9, 1000.0, 1.0, 0.9, 0.999, 1e-8, 0, 0)
    assert L.nbm_apply_update_f32(C.byref(opt), None, None, None, None, None, None) == 1
    assert L.nbm_finalize_step_f32(C.byref(opt), C.byref(net), None, 0, 0, None, None, None, None, None, None) == 1
    assert L.nbm_upload_params(C.byref(net), None, None) == 1
    with pytest.raises(cabi.NbmError):
        cabi.check(L.nbm_assemble_f32(None, None), "nbm_assemble_f32")


def test_balanced_slabs_cover_the_grid_and_relieve_interface_planes():
    """plan.balanced_slabs: contiguous cover of every x plane, at least one plane per device, and fewer planes for the
    devices whose slab holds the interface (sphere of radius 0.5 in [-1, 1]^3: the middle half of the x range)."""
    from jax_dips_b200 import plan as nplan
    tr = mesh.linspace_grid((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), [32, 16, 16])
    phi = lambda pts: pts.norm(dim=1) - 0.5
    assert nplan.balanced_slabs(phi, tr, 1, device="cpu") == [(0, 32)]
    for world in (2, 4, 8):
        slabs = nplan.balanced_slabs(phi, tr, world, device="cpu", list_weight=10.0)
        assert slabs[0][0] == 0 and slabs[-1][1] == 32 and len(slabs) == world
        assert all(slabs[r][1] == slabs[r + 1][0] for r in range(world - 1))
        assert all(b > a for a, b in slabs)
    s8 = nplan.balanced_slabs(phi, tr, 8, device="cpu", list_weight=10.0)
    n = [b - a for a, b in s8]
    assert n[0] > n[3] and n[7] > n[4], n          # interface-free end slabs take more planes
    # without a list weight the planes all cost the same: the reference's equal blocks
    assert nplan.balanced_slabs(phi, tr, 4, device="cpu", list_weight=0.0) == [(0, 8), (8, 16), (16, 24), (24, 32)]
    # more devices than interface-free planes still leaves one plane each
    tiny = mesh.linspace_grid((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), [8, 8, 8])
    s = nplan.balanced_slabs(phi, tiny, 8, device="cpu")
    assert [b - a for a, b in s] == [1] * 8
