"""Secondary configs that share the tile kernels' occupancy settings (A/B aid): Float32 C3, matrix-free C3, Rosenbrock fused."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
import cts_b200 as cts

which = sys.argv[1:] or ["f32", "mf", "ros"]
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = cts.Context(0)
    out = []
    if "f32" in which:
        r = B.run_c3_f32(cts, ctx, stream); out.append(f"f32 {r['ms_per_step']:.3f} ms frac {r['frac']:.3f}")
        torch.cuda.empty_cache()
    if "mf" in which:
        part = cts.SlabPartition(B.NY, 1, 0)
        op, integ, u = B.build_problem(cts, ctx, part, True)
        op.dss(u, None, 0.0)
        ms = B.timed(stream, lambda: cts.step_(integ), 10, warm=3)
        out.append(f"mf {ms:.3f} ms")
        del op, integ, u
        torch.cuda.empty_cache()
    if "ros" in which:
        r = B.run_rosenbrock(cts, ctx, stream, True); out.append(f"ros-fused {r['ms_per_step']:.3f} ms frac {r.get('frac', 0):.3f}")
    if "c4" in which:
        import numpy as np
        for fuse in ("0", "1"):
            os.environ["CTS_JVP_PRECOND_FUSION"] = fuse
            for dt_, nm in ((np.float64, "f64"), (np.float32, "f32")):
                r = B.run_c4(cts, ctx, stream, dt_)
                out.append(f"c4-{nm} fuse={fuse} {r['ms_per_step']:.2f} ms frac {r['frac']:.3f} B/DOF {r['bytes_per_dof_step']:.0f}")
                torch.cuda.empty_cache()
    if "c4mf" in which:
        import numpy as np
        for dt_, nm in ((np.float64, "f64"), (np.float32, "f32")):
            for mf in (False, True):
                r = B.run_c4(cts, ctx, stream, dt_, matrix_free=mf)
                out.append(f"c4-{nm} mf={int(mf)} {r['ms_per_step']:.2f} ms frac {r['frac']:.3f} B/DOF {r['bytes_per_dof_step']:.0f}")
                torch.cuda.empty_cache()
    print(os.environ.get("CTS_LIB_PATH", "default").split("/")[-1], "|", "; ".join(out))
