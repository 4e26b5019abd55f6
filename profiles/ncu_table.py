"""One table row per launch of an .ncu-rep (ncu --set full): time, DRAM bytes, DRAM % of ncu's peak,
launch shape, issue-active %, tensor-pipe %.   python profiles/ncu_table.py REP OUT.md "title" """
import csv
import io
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")

    def get(r, k):
        if k not in hdr:
            return None, ""
        i = hdr.index(k)
        try:
            return float(r[i].replace(",", "")), units[i]
        except ValueError:
            return None, units[i]

    def scale(v, u, want):  # to ms / MB
        if v is None:
            return None
        f = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        return v * f.get(u, 1.0)

    lines = [f"# {title}", "",
             "Per-launch values from `ncu --set full --clock-control none` (cold caches, serialised replays: compare",
             "bytes and shares, not absolute times). dram % = ncu's own DRAM-throughput percentage.", "",
             "| # | kernel | ms | dram read MB | dram write MB | dram % | grid x block | regs | dyn smem KB | issue active % | tensor pipe % | warps active % |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for n, r in enumerate(data):
        v = {k: get(r, k) for k in COLS}
        ms = scale(*v[COLS[0]], "ms")
        rd, wr = scale(*v[COLS[1]], "MB"), scale(*v[COLS[2]], "MB")
        smem = v[COLS[7]][0]
        smem_kb = None if smem is None else smem * {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(v[COLS[7]][1], 1e-3)
        name = r[ki].replace("(anonymous namespace)::", "")[:70]
        f = lambda x, p=1: "" if x is None else f"{x:.{p}f}"
        lines.append(f"| {n} | `{name}` | {f(ms, 4)} | {f(rd)} | {f(wr)} | {f(v[COLS[3]][0])} | "
                     f"{f(v[COLS[4]][0], 0)} x {f(v[COLS[5]][0], 0)} | {f(v[COLS[6]][0], 0)} | {f(smem_kb)} | "
                     f"{f(v[COLS[8]][0])} | {f(v[COLS[9]][0])} | {f(v[COLS[10]][0])} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu summary")
