import os, subprocess, sys
for minb in ["2", "1"]:
    env = dict(os.environ, XCP_DW_MINB=minb)
    out = subprocess.run([sys.executable, "tools/kernel_bench.py", "128", "dw"], env=env, capture_output=True, text=True)
    print("== MINB", minb, out.stderr[-300:] if out.returncode else "")
    for l in out.stdout.splitlines():
        if "affine" in l or "add_full" in l:
            print("  ", l[:64], flush=True)
