"""profiles/scan_kernel_traffic.json from an `ncu --set full` capture of the bench (the file bench.py reads for
roofline.traffic): DRAM bytes read + written by the partition-scan launch of scan_mma_kernel -- the second scan_mma
launch of a step; the first is the coarse scan of the centroid list.
usage: python scripts/ncu_traffic.py gpurun_out/prof_scan_<tag>.ncu-rep profiles/<tag>_ncu_full_scan_mma.json"""
import csv, io, json, os, subprocess, sys
rep, out_summary = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
metrics = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(metrics)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, r):
        if h in metrics:
            x = float(v.replace(",", ""))
            if u == "Mbyte": x *= 1e6
            if u == "Kbyte": x *= 1e3
            if u == "Gbyte": x *= 1e9
            if h == "gpu__time_duration.sum" and u == "ms": x *= 1e3
            if h == "gpu__time_duration.sum" and u == "ns": x /= 1e3
            d[h] = x
    d["kernel"] = r[hdr.index("Kernel Name")]
    launches.append(d)
scan = [l for l in launches if "scan_mma" in l["kernel"]]
part = max(scan, key=lambda l: l["dram__bytes_read.sum"])  # the partition scan streams the index; the coarse scan 2 MB
json.dump({"capture": os.path.basename(rep), "launches": launches}, open(out_summary, "w"), indent=1)
alg = 511012352
t = {"kernel": "scan_mma_kernel<false> (partition scan, C2: 1M x 128, nlist 4096, Q 1024, nprobe 64, k 10)",
     "source": f"{os.path.relpath(out_summary, ROOT)} (ncu --set full --clock-control none, one launch; cold caches)",
     "dram_bytes_read": int(part["dram__bytes_read.sum"]), "dram_bytes_write": int(part["dram__bytes_write.sum"]),
     "dram_bytes_per_launch": int(part["dram__bytes_read.sum"] + part["dram__bytes_write.sum"]),
     "algorithmic_bytes_per_launch": alg,
     "ratio": round((part["dram__bytes_read.sum"] + part["dram__bytes_write.sum"]) / alg, 4),
     "ncu_duration_us": part["gpu__time_duration.sum"]}
json.dump(t, open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json"), "w"), indent=1)
print(json.dumps(t))
