"""Copy the judged summaries of the last GPU measurement pass from gpurun_out/ (scratch) into profiles/round1/."""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles", "round1")
NOISE = ("sm__ops_path", "hmma", ".max", ".min", ".sum.pct", "utccp", "TriageCompute")


def ncu_summary(rep, title, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    summ = subprocess.run(["python", os.path.join(ROOT, "scripts", "ncu_summary.py")], input=raw, capture_output=True,
                          text=True).stdout
    with open(out, "w") as f:
        f.write("# " + title + "\n# source report: gpurun_out/%s (scratch, not committed); metric, unit, value\n"
                % os.path.basename(rep))
        for line in summ.splitlines():
            if not any(n in line for n in NOISE):
                f.write(line + "\n")


def launches(csv_path, out, title):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split("(")[0], []).append(float(r[-1]) / 1e6)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("# " + title + "\n# (cold-cache, serialised: compare SHARES)  kernel, launches, total_ms, avg_ms, share\n")
        for k, v in agg.items():
            f.write("%s, %d, %.3f, %.4f, %.4f\n" % (k, len(v), sum(v), sum(v) / len(v), sum(v) / tot))


if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for name in ("bench_c4_n1", "bench_c3_n1", "bench_c2_n1", "bench_c5_n1", "bench_c4_n1_fp64dmma", "bench_ref_n1"):
        src = os.path.join(SRC, name + ".json")
        if os.path.exists(src) and os.path.getsize(src) > 0:
            shutil.copy(src, os.path.join(DST, name + ".json"))
    shutil.copy(os.path.join(SRC, "launches_c4.csv"), os.path.join(DST, "launches_c4_ncu.csv"))
    launches(os.path.join(SRC, "launches_c4.csv"), os.path.join(DST, "launches_c4_summary.txt"),
             "ncu launch list, bench.py --steps 1 --warmup 1 --e2e-steps 1 (C4, B=8192/GPU, default tri_mode 4), "
             "rollout kernels only")
    cmd = "ncu --set full --clock-control none --import-source on -k regex:%s -s 2 -c 1 (bench.py C4, B=8192, N=5000, n_s=4)"
    for rep, pat, out in (("prof_tri_i8m_c4.ncu-rep", "tri_i8m", "tri_i8m_c4_ncu_full.txt"),
                          ("prof_tri_i8x2_c4.ncu-rep", "tri_i8x2", "tri_i8x2_c4_ncu_full.txt"),
                          ("prof_kstar_i8_c4.ncu-rep", "kstar_i8", "kstar_i8_c4_ncu_full.txt"),
                          ("prof_ellipsoid_c4.ncu-rep", "ellipsoid_step", "ellipsoid_step_c4_ncu_full.txt")):
        if os.path.exists(os.path.join(SRC, rep)):
            ncu_summary(os.path.join(SRC, rep), cmd % pat, os.path.join(DST, out))
    tri = open(os.path.join(DST, "tri_i8m_c4_ncu_full.txt")).read()
    rd = [l for l in tri.splitlines() if l.startswith("dram__bytes_read.sum,")][0].split(",")
    wr = [l for l in tri.splitlines() if l.startswith("dram__bytes_write.sum,")][0].split(",")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    total = float(rd[2]) * scale[rd[1].strip()] + float(wr[2]) * scale[wr[1].strip()]
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(tpath))
    t["C4:tri_mode4"] = int(total)
    json.dump(t, open(tpath, "w"), indent=1)
    print("traffic C4 tri_mode4:", int(total))
