"""Reference-side binding (sketch, needs jax >= 0.4.31 for jax.ffi): registers the XLA-FFI handler of
integration/nbm_xla_ffi.cc and wraps it in a custom_vjp so that `jax.value_and_grad(nbm_loss)` returns the CUDA
step's loss and gradient.  Not importable in this repository's image (jax is not installed); see INTEGRATION.md."""
# jax_dips/solvers/poisson/trainer.py  (reference side; replaces the body of Trainer.loss)
import ctypes, jax, jax.numpy as jnp
_lib = ctypes.CDLL("libnbm_xla_ffi.so")
jax.ffi.register_ffi_target("nbm_loss_grad", jax.ffi.pycapsule(_lib.NbmLossGrad), platform="CUDA")

def _loss_grad(flat_params, plan_addr, P):
    out = jax.ffi.ffi_call("nbm_loss_grad", jax.ShapeDtypeStruct((P + 1,), jnp.float32))(flat_params, plan=plan_addr)
    return out[P], out[:P]

@jax.custom_vjp
def nbm_loss(flat_params, plan_addr, P):            # same value as Trainer.loss(params, points, dx, dy, dz)
    return _loss_grad(flat_params, plan_addr, P)[0]
def _fwd(flat_params, plan_addr, P):
    loss, grad = _loss_grad(flat_params, plan_addr, P)
    return loss, grad
def _bwd(grad, ct):
    return (ct * grad, None, None)
nbm_loss.defvjp(_fwd, _bwd)
# Trainer.update keeps: loss, grads = value_and_grad(nbm_loss)(ravel(params), plan, P); optax as before,
# or drops optax too and calls nbm_apply_update_f32 through a second ffi target.
