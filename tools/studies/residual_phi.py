import sys
sys.path.insert(0, '.')
import numpy as np
import altair_raytracing_b200 as A
N = 4_000_000_000
bands = [(0, 20), (20, 40), (40, 60), (60, 80), (80, 100), (100, 140), (140, 180)]
# phi sectors centred on +x (phi=0), +y (90), -x (180), -y (270): 4-degree bins -> 90 bins; sector = 22-23 bins
sect = {"+x": list(range(79, 90)) + list(range(0, 11)), "+y": list(range(11, 34)), "-x": list(range(34, 56)), "-y": list(range(56, 79))}
with A.Context([0]) as ctx:
    for name, th in (("perposition_170_dir5_0_0", 170.0), ("perposition_163_dir5_0_0", 163.0)):
        k_ref = np.load(f"tests/golden/{name}.npz")["hits"].astype(float).reshape(180, 90)
        c, st = ctx.trace_fluxmap(A.scene(theta_max=th), A.source(), N, A.map_spec(mode=A.MAP_LINE), seed=23)
        exp = 50000.0 * c[0].reshape(180, 90).astype(float) / N
        print(f"== theta_max {th}")
        for sname, cols in sect.items():
            row = []
            for a, b in bands:
                r, e = k_ref[a:b][:, cols].sum(), exp[a:b][:, cols].sum()
                row.append(f"{(r / e - 1) * 100:+5.2f}%({(r - e) / np.sqrt(e):+4.1f})")
            print(f"  phi sector {sname}: " + " ".join(row))
print("bands (theta deg):", [(a / 2, b / 2) for a, b in bands])
