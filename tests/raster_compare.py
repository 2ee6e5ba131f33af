"""Shared parity metric for rasterizer images (north_star: 1e-4 relative fp32).

A pixel is "bad" if |a-b| > 1e-4 * max(1, |b|).  The pipeline has hard thresholds
(alpha < 1/255, T < 1e-4, power > 0, ceil / floor in the tile rectangle) that a 1-ulp
difference in expf or an FMA contraction can flip for isolated pixels (SURVEY.md section 7,
"Branch-boundary parity"), so comparisons carry an explicit outlier budget instead of a max.
"""
import numpy as np

TOL = 1e-4
OBSERVED = {"n": 0, "frac_bad": 0.0, "frac_bad_rel": 0.0, "max_err": 0.0, "max_err_rel": 0.0}  # worst over this session's comparisons


REL_FLOOR = 0.05  # colours live in [0,1]: "relative" below this magnitude is measured against the floor


def compare_images(a, b):
    """frac_bad / max_err: |a-b| / max(1, |b|) (the round-1 criterion, absolute for colours in [0,1]);
    frac_bad_rel / max_err_rel: TRUE relative error |a-b| / max(REL_FLOOR, |b|) -- north_star's "1e-4 relative fp32" read
    per pixel, with a floor so that black pixels (b ~ 0, rounding noise ~1e-8) do not divide by zero."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    diff = np.abs(a - b)
    err = diff / np.maximum(1.0, np.abs(b))
    rel = diff / np.maximum(REL_FLOOR, np.abs(b))
    bad = err > TOL
    bad_rel = rel > TOL
    n = max(err.size, 1)
    OBSERVED["n"] += 1
    for k, v in (("frac_bad", bad.sum() / n), ("frac_bad_rel", bad_rel.sum() / n), ("max_err", err.max() if err.size else 0.0),
                 ("max_err_rel", rel.max() if rel.size else 0.0)):
        OBSERVED[k] = max(OBSERVED[k], float(v))
    return dict(frac_bad=float(bad.sum() / n), max_err=float(err.max()) if err.size else 0.0, n_bad=int(bad.sum()),
                median=float(np.median(err)) if err.size else 0.0, frac_bad_rel=float(bad_rel.sum() / n),
                max_err_rel=float(rel.max()) if rel.size else 0.0, p9999_rel=float(np.quantile(rel, 0.9999)) if rel.size else 0.0)


def load_golden_case(path):
    z = np.load(path)
    inputs = dict(means3D=z["means3D"], opacities=z["opacities"], view=z["view"], proj=z["proj"], campos=z["campos"],
                  W=int(z["W"]), H=int(z["H"]), tan_fovx=float(z["tan_fovx"]), tan_fovy=float(z["tan_fovy"]), bg=z["bg"],
                  sh_degree=int(z["sh_degree"]), scale_modifier=float(z["scale_modifier"]))
    for k_npz, k_arg in (("shs", "shs"), ("colors_precomp", "colors_precomp"), ("scales", "scales"), ("rotations", "rotations"),
                         ("cov3D_precomp", "cov3D_precomp")):
        if k_npz in z.files:
            inputs[k_arg] = z[k_npz]
    return dict(inputs=inputs, color=z["color"], radii=z["radii"], num_rendered=int(z["num_rendered"]), final_T=z["final_T"])
