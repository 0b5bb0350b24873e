"""Copy the judged summaries of the last GPU measurement pass from gpurun_out/ (scratch) into profiles/round2/.

    python scripts/collect_profiles.py [tag]

Bench lines are copied as they are; `.ncu-rep` reports are summarised to text (metric, unit, value) through
scripts/ncu_summary.py; the ncu launch list is kept as CSV plus a per-kernel share table; profiles/traffic.json
(`roofline.traffic` of bench.py) is refreshed from the full capture of the contraction kernel.
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles", "round2")
NOISE = ("sm__ops_path", "hmma", ".max", ".min", ".sum.pct", "utccp", "TriageCompute")


def ncu_summary(rep, title, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py")], input=raw,
                          capture_output=True, text=True).stdout
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


def dram_bytes(summary_path):
    txt = open(summary_path).read()
    rd = [l for l in txt.splitlines() if l.startswith("dram__bytes_read.sum,")][0].split(",")
    wr = [l for l in txt.splitlines() if l.startswith("dram__bytes_write.sum,")][0].split(",")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    return int(float(rd[2]) * scale[rd[1].strip()] + float(wr[2]) * scale[wr[1].strip()])


# gpurun_out/ also holds files of earlier rounds: only what was written after this round began is collected
ROUND_START = 1792231200.0   # 2026-10-17 10:00 UTC


def fresh(path):
    return os.path.getmtime(path) >= ROUND_START


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    os.makedirs(DST, exist_ok=True)
    for fn in sorted(os.listdir(SRC)):
        if (fn.startswith("bench_") and fn.endswith(".json") and os.path.getsize(os.path.join(SRC, fn)) > 0
                and fresh(os.path.join(SRC, fn))):
            shutil.copy(os.path.join(SRC, fn), os.path.join(DST, fn.replace(".json", tag + ".json")))
    lc = os.path.join(SRC, "launches_c4.csv")
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(DST, "launches_c4_ncu%s.csv" % tag))
        launches(lc, os.path.join(DST, "launches_c4_summary%s.txt" % tag),
                 "ncu --metrics gpu__time_duration.sum --clock-control none; bench.py --scaling weak --steps 1 --warmup 2 "
                 "--no-graph (C4 shard, B=8192, automatic digit set), rollout kernels only")
    for c in ("c2", "c3", "c5"):
        lc = os.path.join(SRC, "launches_%s.csv" % c)
        if os.path.exists(lc):
            launches(lc, os.path.join(DST, "launches_%s_summary%s.txt" % (c, tag)),
                     "ncu --metrics gpu__time_duration.sum --clock-control none; bench.py --config %s --steps 1 --no-graph "
                     "(C5: --scaling weak --batch 8192), rollout kernels only" % c.upper())
    lc = os.path.join(SRC, "launches_setup_c4.csv")
    if os.path.exists(lc):
        launches(lc, os.path.join(DST, "launches_setup_c4_summary%s.txt" % tag),
                 "ncu --metrics gpu__time_duration.sum --clock-control none; scripts/profile_setup.py C4 0 with "
                 "SEGP_FACT_I8=1: ONE segp_factorize of the C4 model (4 output dimensions; serialised by ncu) + its probe")
    for fn in sorted(os.listdir(SRC)):
        if fn.startswith("setup_c") and "fact_i8" in fn and fn.endswith(".log"):
            shutil.copy(os.path.join(SRC, fn), os.path.join(DST, fn.replace(".log", tag + ".txt")))
    rep = os.path.join(SRC, "prof_gemm_i8d_c4.ncu-rep")
    if os.path.exists(rep):
        ncu_summary(rep, "ncu --set full --clock-control none --import-source on -k regex:gemm_i8d -s 2 -c 1 "
                    "(scripts/profile_setup.py C4 0: a trailing update of the C4 factorisation)",
                    os.path.join(DST, "gemm_i8d_c4_ncu_full%s.txt" % tag))
    cmd = ("ncu --set full --clock-control none --import-source on -k regex:%s -s 4 -c 1 "
           "(bench.py --scaling weak: C4 shard, B=8192, N=5000, n_s=4, automatic digit set)")
    for rep, pat, out in (("prof_tri_i8m_c4.ncu-rep", "tri_i8m", "tri_i8m_c4_ncu_full%s.txt" % tag),
                          ("prof_kstar_i8_c4.ncu-rep", "kstar_i8", "kstar_i8_c4_ncu_full%s.txt" % tag),
                          ("prof_ellipsoid_c4.ncu-rep", "ellipsoid_step", "ellipsoid_step_c4_ncu_full%s.txt" % tag)):
        if os.path.exists(os.path.join(SRC, rep)):
            ncu_summary(os.path.join(SRC, rep), cmd % pat, os.path.join(DST, out))
    tri = os.path.join(DST, "tri_i8m_c4_ncu_full%s.txt" % tag)
    if os.path.exists(tri):
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        t = json.load(open(tpath)) if os.path.exists(tpath) else {}
        line = [l for l in open(tri).read().splitlines() if l.startswith("launch__grid_size")]
        digits = 4 if "tri_i8m_kernel<2, 1>" in open(os.path.join(SRC, "launches_c4.csv")).read() else 5
        t["C4:tri_mode4:digits%d" % digits] = dram_bytes(tri)
        json.dump(t, open(tpath, "w"), indent=1, sort_keys=True)
        print("traffic C4 tri_mode4 digits", digits, t["C4:tri_mode4:digits%d" % digits], line)
