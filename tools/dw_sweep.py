import os, subprocess, sys
for halo, st in [(400, 0), (272, 0), (200, 0), (140, 0), (100, 0), (400, 2), (272, 2)]:
    env = dict(os.environ, XCP_DW_HALO=str(halo), XCP_DW_STAGES=str(st))
    out = subprocess.run([sys.executable, "tools/kernel_bench.py", "128", "dw"], env=env, capture_output=True, text=True).stdout
    sel = [l for l in out.splitlines() if "dw_fwd affine" in l and ("19x19" in l or "147x147x128" in l or "37x37x728" in l)]
    print("halo", halo, "stages", st, " | ".join(l.split()[2] + " " + l.split()[3] + "us" for l in sel), flush=True)
