"""Statistics of SURVEY.md section 8(d)'s full-render parity bounds, shared by tests/test_gpu_render.py and scripts/stat_probe.py.

Two renders of the same frame at equal spp (dicts with the structured planes Dd, Ds, Id, Is and the G-buffer) are compared
  * per pixel and plane through z = |mean_a - mean_b| / sqrt((Var_a + Var_b) / n): the Var planes hold the per-sample second
    moment minus the squared mean (calcVar, src/render.cpp:510-516), n is the number of samples that plane received
    (spp_direct, or spp_indirect x 16 on glass, src/render.cpp:498-501);
  * through the trimmed relative MSE of tests/test_gpu_render.py (the CPU-vs-CPU value is the noise floor);
  * through the energy of the mean image (sum of the four planes over the frame).
"""
import numpy as np

PLANES = ("Dd", "Ds", "Id", "Is")


def sample_counts(gbuffer, args):
    spp_d = int(np.float32(args.spp) * np.float32(args.P_Direct))
    base = args.spp - spp_d
    hit = ~np.isnan(gbuffer["position"][:, 0])
    emissive = np.abs(gbuffer["emission"]).sum(1) > 0
    glass = gbuffer["opacity"] <= np.float32(1.0) - np.float32(1e-4)
    n_d = np.where(hit & ~emissive, spp_d, 0)
    n_i = np.where(hit & ~emissive, base * np.where(glass, 16, 1), 0)
    return {"Dd": n_d, "Ds": n_d, "Id": n_i, "Is": n_i}


def z_scores(a, b, args):
    """per plane: |z| of every pixel that received samples"""
    n = sample_counts(b["gbuffer"], args)
    out = {}
    for k in PLANES:
        ra, rb = a[k]["radiance"].astype(np.float64), b[k]["radiance"].astype(np.float64)
        va, vb = a[k]["Var"].astype(np.float64), b[k]["Var"].astype(np.float64)
        ok = n[k] > 0
        d = np.sqrt(((ra - rb) ** 2).sum(1))
        scale = np.maximum(np.abs(ra).max(1), np.abs(rb).max(1))
        se = np.sqrt((va + vb) / np.maximum(n[k], 1) + (1e-3 * scale) ** 2 + 1e-14)
        out[k] = (d / se)[ok]
    return out


def rel_mse(a, b, trim=0.01):
    e = np.sort(((a - b) ** 2).sum(1) / ((b ** 2).sum(1) + 1e-2))
    return float(e[: max(1, int(len(e) * (1.0 - trim)))].mean())


def energy(r):
    return float(sum(r[k]["radiance"].astype(np.float64).sum() for k in PLANES))


def compare_renders(a, b, args):
    z = z_scores(a, b, args)
    allz = np.concatenate([z[k] for k in PLANES]) if any(len(z[k]) for k in PLANES) else np.zeros(1)
    res = {"mean_abs_z": float(allz.mean()), "p_abs_z_gt4": float((allz > 4.0).mean()), "energy_ratio": energy(a) / max(energy(b), 1e-30),
           "planes": {}}
    for k in PLANES:
        ra, rb = a[k]["radiance"].astype(np.float64), b[k]["radiance"].astype(np.float64)
        res["planes"][k] = {"mean_abs_z": float(z[k].mean()) if len(z[k]) else 0.0, "p_abs_z_gt4": float((z[k] > 4.0).mean()) if len(z[k]) else 0.0,
                            "rel_mse": rel_mse(ra, rb), "energy_ratio": float(ra.sum() / max(rb.sum(), 1e-30))}
    return res
