"""Shared comparison helpers for the parity tests."""
import numpy as np


def bpp_of(lik_y, lik_z, num_pixels):
    return (np.log(lik_y.astype(np.float64)).sum() + np.log(lik_z.astype(np.float64)).sum()) / (-np.log(2) * num_pixels)


def compare_forward(got, g, num_pixels, *, cont_tol=1e-4, flip_frac=1e-3, yhat_frac=1e-2, xhat_max=1e-2,
                    xhat_psnr=60.0, bpp_rel=1e-3, lik_tol=1e-3):
    """`got`/`g`: dicts of numpy arrays (g = golden from the reference).

    Continuous tensors upstream of the quantiser must agree to `cont_tol`.
    Downstream of it the comparison is statistical: a 1e-6 wobble in mu can move
    a value across a round-half tie, which changes that symbol by one and then
    legitimately perturbs later slices (SURVEY section 7 "bit-exactness").  Integer
    outputs are asserted bit-exact *given identical inputs* in the op-level tests.
    """
    for k in ("y", "z", "latent_means", "latent_scales"):
        if k in got:
            np.testing.assert_allclose(got[k], g[k], rtol=cont_tol, atol=cont_tol, err_msg=k)
    for k in ("symbols", "indexes"):
        if k in got and got[k] is not None:
            frac = (got[k] != g[k]).mean()
            assert frac <= flip_frac, (k, frac)
    if "y_hat" in got:
        frac = (np.abs(got["y_hat"] - g["y_hat"]) > 10 * cont_tol).mean()
        assert frac <= yhat_frac, ("y_hat", frac)
    dx = got["x_hat"].astype(np.float64) - g["x_hat"]
    assert np.abs(dx).max() <= xhat_max, ("x_hat max abs", np.abs(dx).max())
    psnr = -10 * np.log10((dx ** 2).mean() + 1e-30)
    assert psnr >= xhat_psnr, ("x_hat psnr vs reference", psnr)
    np.testing.assert_allclose(got["lik_z"], g["lik_z"], rtol=lik_tol, atol=1e-9)
    bad = np.abs(got["lik_y"] - g["lik_y"]) > 1e-4 + lik_tol * g["lik_y"]
    assert bad.mean() <= max(yhat_frac, 1e-3), ("lik_y", bad.mean())
    b_got, b_ref = bpp_of(got["lik_y"], got["lik_z"], num_pixels), bpp_of(g["lik_y"], g["lik_z"], num_pixels)
    assert abs(b_got - b_ref) <= bpp_rel * b_ref, ("bpp", b_got, b_ref)
    return {"x_hat_max": float(np.abs(dx).max()), "x_hat_psnr": float(psnr), "bpp": float(b_got), "bpp_ref": float(b_ref)}
