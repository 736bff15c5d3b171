"""Distils `ncu --set full` captures (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into profiles/ncu_traffic.json:
per kernel family the per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured
launches), duration and the utilisation figures bench.py's `roofline` / `roofline_ar` quote as `traffic`.

    python tools/ncu_traffic.py <tag> <kernel family>=<file.ncu-rep> [...]      # merges into profiles/ncu_traffic.json

and writes profiles/<tag>_<family>_ncu_summary.txt (one line per captured launch).  Run in the build container."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
              "nsecond": 1e-3}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    res = []
    for row in r[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")], "grid": row[hdr.index("Grid Size")], "block": row[hdr.index("Block Size")]}
        for k, name in KEYS.items():
            if name in hdr:
                i = hdr.index(name)
                try:
                    d[k] = float(row[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0)
                except ValueError:
                    d[k] = None
        res.append(d)
    return res


def main():
    tag, pairs = sys.argv[1], [a.split("=", 1) for a in sys.argv[2:]]
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    path = ROOT / "profiles" / "ncu_traffic.json"
    table = json.loads(path.read_text()) if path.exists() else {}
    for family, rep in pairs:
        rows = rows_of(rep)
        lines = []
        for d in rows:
            lines.append(" ".join(f"{k}={d[k]:.4g}" if isinstance(d[k], float) else f"{k}={d[k]}" for k in d))
        summary = ROOT / "profiles" / f"{tag}_{family}_ncu_summary.txt"
        summary.write_text(f"# {rep} ({len(rows)} launches), ncu --set full --clock-control none; bytes are per launch\n" + "\n".join(lines) + "\n")
        n = len(rows)
        table[family] = {
            "dram_bytes_per_launch": sum((d["dram_read"] or 0) + (d["dram_write"] or 0) for d in rows) / n,
            "duration_us_under_ncu": sum(d["duration_us"] for d in rows) / n,
            "dram_pct_of_peak": sum(d["dram_pct"] for d in rows) / n,
            "tensor_pipe_active_pct": sum(d["tensor_pct"] or 0 for d in rows) / n,
            "launches_captured": n,
            "source": f"profiles/{summary.name} (ncu --set full, capture tag {tag}, repo at {commit})",
        }
    path.write_text(json.dumps(table, indent=1) + "\n")
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
