"""Turns the raw evidence of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/:
launch list per kernel (time share + DRAM bytes), key ncu metrics of the --set full captures, roofline inputs.
Runs without a GPU:  python tools/make_profile_summary.py r01"""
import collections, csv, json, os, re, shutil, subprocess, sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")
os.makedirs(DST, exist_ok=True)

def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void b200cv::<unnamed>::", "").replace("b200cv::<unnamed>::", "").replace("void ", "")

# ---- launch list ---------------------------------------------------------------------------------------------
roof = None
if os.path.exists(os.path.join(SRC, f"launches_{R}.csv")):
    rows = list(csv.reader(open(os.path.join(SRC, f"launches_{R}.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, mn, mv, gs, ident = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("ID")
    launch = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        d = launch.setdefault(r[ident], {"name": short(r[kn]), "grid": r[gs]})
        d[r[mn]] = float(r[mv].replace(",", ""))
    agg = collections.OrderedDict()
    for d in launch.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e3       # ns -> us
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(DST, f"launches_{R}_summary.txt"), "w") as f:
        f.write(f"# ncu launch list of ONE eager Darknet-53 416^2 bs64 training step (bench.py --ncu-window), {len(launch)} launches,\n"
                f"# sum of gpu__time_duration = {tot/1e3:.2f} ms (cold-cache, serialised: the SHARE is what compares with the bench).\n"
                f"# units: us, MB per step over all launches of the kernel\n")
        f.write(f"{'kernel':58s} {'n':>5s} {'us':>10s} {'share':>7s} {'us/launch':>10s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s}\n")
        for k, (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} {n:5d} {us:10.1f} {100*us/tot:6.1f}% {us/n:10.2f} {rd/1e6:11.1f} {wr/1e6:11.1f}\n")
    conv = [a for k, a in agg.items() if k.startswith(("igemm_kernel", "wgrad_kernel", "conv_image_"))]
    roof = {"round": R, "conv_launches_per_step": sum(a[0] for a in conv), "conv_us_ncu": sum(a[1] for a in conv),
            "conv_share_of_step_ncu": sum(a[1] for a in conv) / tot,
            "conv_dram_bytes_per_step": sum(a[2] + a[3] for a in conv),
            "note": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) summed over the igemm / wgrad / conv_image launches of one "
                    "eager step, from the ncu launch-list pass"}
    json.dump(roof, open(os.path.join(DST, f"roofline_{R}.json"), "w"), indent=1)


# ---- --set full captures ---------------------------------------------------------------------------------------
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum"]
for rep in sorted(os.listdir(SRC)):
    if not rep.endswith(f"_{R}.ncu-rep"):
        continue
    out = subprocess.run(["ncu", "-i", os.path.join(SRC, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    if len(rr) < 3:
        continue
    hdr, units = rr[0], rr[1]
    with open(os.path.join(DST, "ncu_" + rep.replace(".ncu-rep", ".txt")), "w") as f:
        f.write(f"# key metrics of gpurun_out/{rep} (ncu --set full --clock-control none --import-source on)\n")
        for r in rr[2:]:
            f.write(f"== {short(r[hdr.index('Kernel Name')])[:90]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"   {k:72s} {r[i]:>18s} {units[i]}\n")
    # SASS evidence of tcgen05 / TMA for the conv kernels
    if rep.startswith(("igemm", "wgrad", "dgrad", "optim", "detect", "image")):
        src = subprocess.run(["ncu", "-i", os.path.join(SRC, rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
        ops = collections.Counter(m for m in re.findall(r"\b(UTCHMMA|HMMA[.\w]*|LDSM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAREDG[.\w]*|UTCBAR|LDTM[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*)", src))
        with open(os.path.join(DST, "ncu_" + rep.replace(".ncu-rep", ".txt")), "a") as f:
            f.write("# SASS mnemonics (static count over the captured kernels): " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items())) + "\n")

for name in (f"bench_{R}.json", f"layer_times_{R}.txt", f"yolo_loss_timing_{R}.log", f"pipeline_{R}.json", f"optim_{R}.json",
             f"determinism_{R}.txt", "tma_bench.txt", "harness_bench_h.txt", "dbg_sweep.txt"):
    p = os.path.join(SRC, name)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(DST, name if R in name else name.replace(".txt", f"_{R}.txt")))
if roof is not None:
    print(open(os.path.join(DST, f"launches_{R}_summary.txt")).read())
    print(json.dumps(roof))
