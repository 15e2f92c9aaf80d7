"""Generate golden vectors by running the REFERENCE'S OWN SOURCE FILES (read in place from
/root/reference, never copied) through the numpy stand-in for jax in oracle/jax_shim.

    python oracle/make_golden.py            # writes tests/golden/*.npz   (build container only)

What runs from the reference, unmodified:
  jax_dips/domain/mesh.py, jax_dips/domain/interpolate.py (multilinear_interpolation,
  nonoscillatory_quadratic_interpolation_per_point, add_ghost_layer_3d),
  jax_dips/geometry/level_set.py (perturb_level_set_fn),
  jax_dips/geometry/geometric_integrations_per_point.py (all), and
  jax_dips/solvers/poisson/discretization.py (Discretization: get_regression_coeffs_at_point,
  get_u_mp_by_regression_at_point_fn, compute_Ax_and_b_preconditioned_fn).
The three hooks the reference's Trainer adds on top of Discretization (solution_at_point_fn,
evaluate_solution_fn, precond_fn; trainer.py:836-854) need haiku, which is not installed: they are
supplied here as the plain `hk.Linear`/tanh stack of nn/mlp/MLP.py:93-139 with the disabled
preconditioner (trainer.py:246-253).

Each fixture stores the inputs (grids, phi samples, points, cell size, parameter vector) and the
reference's outputs in float32 mode AND in x64 mode.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
if not os.path.isdir(REF):
    raise SystemExit("the reference tree is only available in the build container")
sys.path.insert(0, os.path.join(HERE, "jax_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import jax  # the stand-in  # noqa: E402
from jax import numpy as jnp  # noqa: E402
from jax_dips.domain import interpolate, mesh  # noqa: E402
from jax_dips.geometry import geometric_integrations_per_point as gipp  # noqa: E402
from jax_dips.geometry import level_set  # noqa: E402
from jax_dips.solvers.poisson.discretization import Discretization  # noqa: E402
from jax_dips.solvers.simulation_states import PoissonSimStateFn  # noqa: E402

from jax_dips_b200 import numpy as tnp  # noqa: E402
from jax_dips_b200 import problems  # noqa: E402
from oracle import nbm_oracle as O  # noqa: E402  (only for the shared initial parameter vector)


def torch_batched(fn, dtype):
    """coefficient callables are INPUTS: the same torch definitions the product and the oracle use"""
    v = tnp.vmap(fn)
    tdt = torch.float64 if dtype == np.float64 else torch.float32

    def f(R):
        R = np.asarray(R, dtype=dtype).reshape(-1, 3)
        return jnp.Arr(v(torch.from_numpy(np.ascontiguousarray(R)).to(tdt)).numpy().astype(dtype))
    return f


def unpack(flat, L, H, off):
    layers, fan_in = [], 3
    for _ in range(L):
        W = flat[off: off + fan_in * H].reshape(fan_in, H); off += fan_in * H
        b = flat[off: off + H]; off += H
        layers.append((W, b)); fan_in = H
    W = flat[off: off + fan_in].reshape(fan_in, 1); off += fan_in
    layers.append((W, flat[off: off + 1]))
    return layers


def nonlinear_callable(op):
    """the reference takes the nonlinear operator as a callable on u (discretization.py:369;
    examples/biomolecules/coefficients.py:126-131): build it from the product's named operator"""
    if op is None or getattr(op, "kind", 0) == 0:
        return lambda u: 0.0
    coef = float(op.coef)
    return lambda u: coef * jnp.sinh(u)


class Hooked(Discretization):
    """Discretization + the three Trainer hooks (trainer.py:836-854), network = MLP.py:93-139"""

    # post-processing hooks the Trainer defines (trainer.py:914-957); not on the training path
    compute_normal_gradient_solution_mp_on_interface_neural_network = None
    compute_gradient_solution_mp_neural_network = None
    compute_normal_gradient_solution_on_interface_neural_network = None
    compute_gradient_solution_neural_network = None

    def set_net(self, flat, shape):
        self.flat, self.shape = flat, shape

    def _mlp(self, r, L, H, off):
        h = r
        layers = unpack(self.flat, L, H, off)
        for (W, b) in layers[:-1]:
            h = jnp.tanh(h @ W + b)
        W, b = layers[-1]
        return h @ W + b

    def _forward(self, r, phi):
        s = self.shape
        return jnp.where(phi >= 0, self._mlp(r, s.Lp, s.Hp, 0), self._mlp(r, s.Lm, s.Hm, s.n_p))

    def solution_at_point_fn(self, params, r_point, phi_point):
        return self._forward(r_point, phi_point).reshape()

    def evaluate_solution_fn(self, params, R_flat):
        phi_flat = self.phi_interp_fn(R_flat)
        return jax.vmap(self._forward, (0, 0))(R_flat, phi_flat)

    def precond_fn(self, params, lhs_rhs):
        return 1.0


class HookedPrecond(Hooked):
    """+ the reference's learned preconditioner: the REFERENCE'S OWN flax module (jax_dips/nn/preconditioner.py:10-35,
    run through the flax stand-in oracle/jax_shim/flax) behind the Trainer's hook (trainer.py:846-847)"""

    def set_precond(self, widths, scaling_coeff):
        from jax_dips.nn.preconditioner import Preconditioner
        self.precond = Preconditioner(Ds=list(widths), out_dim=1, scaling_coeff=scaling_coeff)   # trainer.py:230-234

    def precond_fn(self, params, lhs_rhs):
        return self.precond.apply(params["preconditioner"], lhs_rhs)


def precond_tree(flat, widths, n_in=26):
    """flat [Dense_0.kernel (in,out) row-major, Dense_0.bias, Dense_1.kernel, ...] -> flax's variable tree"""
    tree, off, fan_in = {}, 0, n_in
    for i, d in enumerate(list(widths) + [1]):
        k = flat[off: off + fan_in * d].reshape(fan_in, d); off += fan_in * d
        b = flat[off: off + d]; off += d
        tree[f"Dense_{i}"] = {"kernel": jnp.Arr(k), "bias": jnp.Arr(b)}
        fan_in = d
    assert off == flat.size
    return {"params": tree}


PRECOND = ((8, 4), 1.0)     # examples/benchmark_LPBE/conf/lpbe.yaml:62-67


def precond_flat(dtype):
    pp = O.init_precond_params(O.PrecondShape(*PRECOND), seed=1, dtype=torch.float64)
    pp[:26 * 8] *= 20.0      # cell coefficients are O(h): larger first-layer weights make P vary visibly between cells
    return pp.numpy().astype(dtype)


def run_case(name, problem, n_tr, n_lvl, interp, point_idx, zoom, dtype, precond=False):
    x64 = dtype == np.float64
    jax.config.update("jax_enable_x64", x64)
    lo, hi = problem.box
    init_mesh_fn, _ = mesh.construct(3)
    ax = lambda n, a: jnp.linspace(lo[a], hi[a], n, dtype=jnp.float32)   # the drivers build float32 grids
    # (the training mesh only supplies dx = x[1] - x[0], mesh.py:121-153: not built, a 256^3 `R` is 200 MB)
    tr_d = [ax(n_tr, a)[1] - ax(n_tr, a)[0] for a in range(3)]
    lv = init_mesh_fn(ax(n_lvl, 0), ax(n_lvl, 1), ax(n_lvl, 2))
    phi_grid = tnp.vmap(problem.phi_fn)(torch.from_numpy(np.asarray(lv.R, dtype=np.float32))).numpy()
    # level set = the reference's grid interpolant of the float32 samples (+ perturbation)
    if interp == "trilinear":
        base = interpolate.multilinear_interpolation(jnp.array(phi_grid), lv)
    else:
        single = interpolate.nonoscillatory_quadratic_interpolation_per_point(jnp.array(phi_grid), lv)
        base = jax.vmap(single)
    phi_fn = level_set.perturb_level_set_fn(base)
    b = lambda fn: torch_batched(fn, dtype)
    fns = PoissonSimStateFn(b(problem.initial_value_fn), b(problem.dirichlet_bc_fn), phi_fn, b(problem.mu_m_fn),
                            b(problem.mu_p_fn), b(problem.k_m_fn), b(problem.k_p_fn), b(problem.f_m_fn),
                            b(problem.f_p_fn), b(problem.alpha_fn), b(problem.beta_fn),
                            nonlinear_callable(problem.nonlinear_op_m), nonlinear_callable(problem.nonlinear_op_p))
    D = (HookedPrecond if precond else Hooked)(lv, None, fns, precondition=1, algorithm=0)
    shape = O.NetShape()
    flat = O.init_params(shape, seed=7, dtype=torch.float64).numpy().astype(dtype)
    D.set_net(jnp.Arr(flat), shape)
    tree = None
    if precond:
        D.set_precond(*PRECOND)
        pc_flat = precond_flat(dtype)
        tree = {"preconditioner": precond_tree(pc_flat, PRECOND[0])}
    d = [dtype(np.float32(v) * np.float32(0.5 ** zoom)) for v in tr_d]
    pts = grid_points(lo, hi, n_tr, point_idx).astype(dtype)
    out = {"lhs_rhs": [], "coeffs": [], "flag": [], "beta_gamma": [], "u_mp": [], "zeta_gamma": []}
    if precond:
        out["precond"] = []
    for p in pts:
        p = jnp.Arr(p)
        out["lhs_rhs"].append(np.asarray(D.compute_Ax_and_b_fn(tree, p, *d)).reshape(2))
        if precond:
            c26 = D.compute_face_centroids_values_plus_minus_at_point(p, *d)
            out["precond"].append(float(np.asarray(D.precond_fn(tree, c26)).reshape(())))
        out["coeffs"].append(np.asarray(D.compute_face_centroids_values_plus_minus_at_point(p, *d)))
        out["flag"].append(float(D.is_cell_crossed_by_interface(p, *d)))
        out["beta_gamma"].append(float(np.asarray(D.beta_integrate_over_interface_at_point(p, *d))))
        out["u_mp"].append(np.asarray(D.u_mp_fn(None, *d, p)).reshape(2))
        rc = D.get_regression_coeffs_at_point(p, *d)
        out["zeta_gamma"].append(np.concatenate([np.asarray(v).reshape(-1) for v in rc[1:]]))
    res = {k: np.asarray(v) for k, v in out.items()}
    res.update(points=pts, d=np.asarray(d), params=flat, phi_grid=phi_grid.astype(np.float32))
    if precond:
        res["pc_params"] = pc_flat
    return res


def grid_points(lo, hi, n_tr, idx):
    """coordinates of flat (z-fastest) indices of the float32 linspace training grid (mesh.py:121-153)"""
    g = [np.linspace(lo[a], hi[a], n_tr).astype(np.float32) for a in range(3)]
    idx = np.asarray(idx, dtype=np.int64)
    return np.column_stack((g[0][idx // (n_tr * n_tr)], g[1][(idx // n_tr) % n_tr], g[2][idx % n_tr]))


def choose_points(problem, n_tr, n_lvl, n_bulk=24, n_near=60, seed=0, max_candidates=1 << 21):
    """boundary rows, bulk rows on both sides, and every kind of interface-adjacent row.  Grids above 128^3 are
    searched through a random subset of their nodes (the analytic level set is evaluated in chunks)."""
    lo, hi = problem.box
    rng = np.random.default_rng(seed)
    n = n_tr ** 3
    cand = np.arange(n) if n <= max_candidates else np.unique(rng.integers(0, n, max_candidates))
    R = grid_points(lo, hi, n_tr, cand)
    v = tnp.vmap(problem.phi_fn)
    phi = np.concatenate([v(torch.from_numpy(R[i: i + 100000])).numpy() for i in range(0, len(R), 100000)])
    g = np.linspace(lo[0], hi[0], n_tr, dtype=np.float32)
    h = g[1] - g[0]
    near = cand[np.abs(phi) < 1.8 * h]
    far = cand[np.abs(phi) >= 1.8 * h]
    bnd = cand[(np.abs(R) >= hi[0] - 1e-6).any(axis=1)]
    sel = np.concatenate((rng.choice(near, min(n_near, near.size), replace=False),
                          rng.choice(far, n_bulk, replace=False), rng.choice(bnd, 8, replace=False)))
    return np.unique(sel)


CASES = [
    # name, problem, n_tr, n_lvl, interp, zoom, near-interface points, learned preconditioner
    ("sphere_tri_z0", "sphere", 16, 32, "trilinear", 0, 60, False),
    ("sphere_tri_z1", "sphere", 16, 32, "trilinear", 1, 60, False),
    ("star_tri_z0", "star", 16, 32, "trilinear", 0, 60, False),
    ("sphere_quad_z0", "sphere", 12, 24, "quadratic", 0, 30, False),
    ("sphere_reaction_tri_z0", "sphere_reaction", 16, 32, "trilinear", 0, 60, False),   # k != 0 and N(u) = c sinh(u) on both sides
    # the learned preconditioner of lpbe.yaml through the reference's own flax module
    ("sphere_precond_tri_z0", "sphere", 16, 32, "trilinear", 0, 60, True),
    # >= 200 crossed cells
    ("sphere_dense_tri_z0", "sphere", 32, 32, "trilinear", 0, 520, False),
    ("star_dense_tri_z0", "star", 32, 32, "trilinear", 0, 520, False),
    # BASELINE.json's other geometries at their own training spacing (level set sampled on a 64^3 lvl grid to keep
    # the fixture small): 64 stars with variable mu, the dragon-like blob union through the quadratic interpolant,
    # the multi-atom Poisson-Boltzmann surface with sinh on the plus side
    ("stars_tri_z0", "stars", 64, 64, "trilinear", 0, 80, False),
    ("dragon_quad_z0", "dragon_like", 128, 64, "quadratic", 0, 80, False),
    ("pb_tri_z0", "poisson_boltzmann", 256, 64, "trilinear", 0, 80, False),
]


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[1:])
    for name, pname, n_tr, n_lvl, interp, zoom, n_near, precond in CASES:
        if only and name not in only:
            continue
        P = problems.PROBLEMS[pname]()
        idx = choose_points(P, n_tr, n_lvl, n_near=n_near)
        blob = {"point_idx": idx, "n_tr": n_tr, "n_lvl": n_lvl, "zoom": zoom}
        for tag, dt in (("f32", np.float32), ("f64", np.float64)):
            r = run_case(name, P, n_tr, n_lvl, interp, idx, zoom, dt, precond=precond)
            for k, v in r.items():
                blob[f"{tag}_{k}"] = v
            print(name, tag, "points", len(idx), "crossed", int((r["flag"] == 0).sum()),
                  "lhs range", float(r["lhs_rhs"][:, 0].min()), float(r["lhs_rhs"][:, 0].max()), flush=True)
        np.savez_compressed(os.path.join(outdir, f"{name}.npz"), **blob)


if __name__ == "__main__":
    main()
