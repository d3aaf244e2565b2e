"""Summarise an `ncu --set full` report (run here, no GPU needed): one row per captured launch with duration, DRAM bytes
and % of peak, tensor pipe %, XU %, issue-slot %, warps active, registers, L2 hit rate.
Usage: python tools/ncu_summary.py gpurun_out/r2_flash_attn.ncu-rep [label ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us", "dur_us"),
    ("dram__bytes_read.sum", None, "dram_rd_MB"),
    ("dram__bytes_write.sum", None, "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "%", "tensor_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "%", "xu_pct"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "%", "tmem_pct"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "%", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%", "warps_pct"),
    ("launch__registers_per_thread", None, "regs"),
    ("lts__t_sector_hit_rate.pct", "%", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%", "sm_pct"),
    ("smsp__cycles_active.avg", None, "cycles"),
]
UNIT_TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
UNIT_TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[hdr], rows[hdr + 1], rows[hdr + 2:]


def main():
    path, labels = sys.argv[1], sys.argv[2:]
    names, units, data = load(path)

    def col(metric):
        for i, n in enumerate(names):          # exact name first (the Triage* section repeats some metrics)
            if n == metric:
                return i
        for i, n in enumerate(names):
            if n.endswith("." + metric):
                return i
        return None

    cols = [(col(m), key) for m, _, key in WANT]
    kcol = names.index("Kernel Name")
    print("| # | kernel | " + " | ".join(k for _, k in cols if _ is not None) + " |")
    print("|---|---|" + "---:|" * sum(1 for c, _ in cols if c is not None))
    for n, r in enumerate(data):
        if not r or len(r) <= kcol:
            continue
        vals = []
        for c, key in cols:
            if c is None:
                continue
            v, u = r[c], units[c]
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                vals.append(v)
                continue
            if key.endswith("_MB"):
                f *= UNIT_TO_MB.get(u, 1e-6)
            if key == "dur_us":
                f *= UNIT_TO_US.get(u, 1.0)
            vals.append(f"{f:.1f}" if abs(f) < 1e5 else f"{f:.0f}")
        name = r[kcol].replace("void ", "").split("(")[0]
        lab = f" — {labels[n]}" if n < len(labels) else ""
        print(f"| {n} | `{name}`{lab} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
