#!/usr/bin/env python
"""Copy the round-2 measurement artefacts from gpurun_out/ (scratch) into profiles/ (tracked) under their documented names."""
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def cp(src, dst):
    src = os.path.join(G, src)
    if os.path.exists(src) and os.path.getsize(src) > 0:
        shutil.copy(src, os.path.join(P, dst))
        print("ok  ", dst)
    else:
        print("miss", src)


def last_json_line(src, dst):
    src = os.path.join(G, src)
    if not os.path.exists(src):
        print("miss", src)
        return
    lines = [l for l in open(src).read().strip().splitlines() if l.startswith("{")]
    if lines:
        json.dump(json.loads(lines[-1]), open(os.path.join(P, dst), "w"), indent=1)
        print("ok  ", dst)


for c in ("C2", "C2_skip", "C3", "C4", "C5"):
    for job in ("r2z", "r2j", "r2g", "r2f", "r2e"):   # newest job that has the line
        if os.path.exists(os.path.join(G, job, f"bench_{c}.json")) and os.path.getsize(os.path.join(G, job, f"bench_{c}.json")) > 0:
            last_json_line(f"{job}/bench_{c}.json", f"bench_r02_{c}.json")
            break
last_json_line(next((f"{j}/bench_ref_C2.json" for j in ("r2z", "r2j", "r2e") if os.path.exists(os.path.join(G, j, "bench_ref_C2.json"))), "r2e/bench_ref_C2.json"), "bench_r02_ref_C2.json")
for n in (2, 8):
    for f in glob.glob(os.path.join(G, f"r2m{n}", "bench_*.json")):
        last_json_line(os.path.relpath(f, G), f"bench_r02_N{n}_" + os.path.basename(f)[6:])
    cp(f"r2m{n}/dp_check.log", f"dp_check_r02_N{n}.txt")
ab = []
for d in ("r2c", "r2d", "r2e", "r2f", "r2g", "r2h", "r2i", "r2o", "r2s"):
    f = os.path.join(G, d, "scan.txt" if d in ("r2h", "r2i", "r2o") else "ab.txt")
    if os.path.exists(f):
        ab.append(f"##### job {d} (one box per job; compare lines within a job only)\n" + open(f).read())
if ab:
    open(os.path.join(P, "ab_r02.txt"), "w").write("\n".join(ab))
    print("ok   ab_r02.txt")
cp("r2r/launches_C2.txt", "launches_r02_C2.txt")
cp("r2z/calls_C2.txt", "calls_r02_C2.txt")
cp("r2p/calls_C2.txt", "calls_r02_C2_before_helper_kernel_fixes.txt")
cp("r2r/calls_C3.txt", "calls_r02_C3.txt")
cp("r2p/calls_C3.txt", "calls_r02_C3_before_vectorised_data_movement.txt")
cp("r2p/scan_R512.txt", "scan_r02_R512_before_48_row_slabs.txt")
cp("r2z/calls_C4.txt", "calls_r02_C4.txt")
cp("r2z/calls_C5.txt", "calls_r02_C5.txt")
cp("r2h/calls_C5.txt", "calls_r02_C5_before_cast_and_slab_fixes.txt")
cp("r2z/tests_all.log", "tests_gpu_all_r02.txt")
cp("r2z/scan_R512.txt", "scan_r02_R512.txt")
cp("r2m/autocast_vgg.txt", "autocast_context_vgg_r02.txt")
cp("r2m/autocast_dcgan.txt", "autocast_context_dcgan_r02.txt")
for rep, out, work in (("r2p/conv_full.ncu-rep", "ncu_conv_r02.txt", ["conv_gemm_kernel=0.503", "convt4_kernel=0.515"]),
                       ("r2z/lstm_full.ncu-rep", "ncu_lstm_r02.txt", [])):
    src = os.path.join(G, rep)
    if os.path.exists(src):
        cmd = [sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), src] + (["--work"] + work if work else [])
        txt = subprocess.run(cmd, capture_output=True, text=True).stdout
        open(os.path.join(P, out), "w").write(f"# ncu --set full --clock-control none ({rep}), summarised by tools/summarize_ncu.py\n\n" + txt)
        print("ok  ", out)
    else:
        print("miss", src)
san = os.path.join(G, "r2san")
if os.path.isdir(san):
    with open(os.path.join(P, "sanitizer_r02.txt"), "w") as f:
        f.write("# compute-sanitizer over the kernel-level GPU tests (tools/gpu_job_san.sh)\n")
        if os.path.exists(os.path.join(san, "rc.txt")):
            f.write(open(os.path.join(san, "rc.txt")).read() + "\n")
        for tool in ("synccheck", "racecheck", "memcheck"):
            p = os.path.join(san, tool + ".log")
            if os.path.exists(p):
                lines = open(p).read().splitlines()
                f.write(f"\n## {tool}: last lines\n" + "\n".join(lines[-25:]) + "\n")
    print("ok   sanitizer_r02.txt")
for d in ("r2e", "r2d"):
    src = os.path.join(G, d, "tests_measured.log")
    if os.path.exists(src):
        keep = [l[:400] for l in open(src).read().splitlines() if "worst gradient cosines" in l or "quantiles" in l or "two-step" in l or " passed" in l or " failed" in l]
        open(os.path.join(P, "tests_measured_r02.txt"), "w").write("# pytest tests/test_measured_gpu.py -s (B200): gradient cosines of the bf16 mode vs the fp32 oracle at the measured shapes\n" + "\n".join(keep) + "\n")
        print("ok   tests_measured_r02.txt")
        break
