"""Top stall sites of a kernel from an ncu report's source page (SASS view): address, samples, main stall reason."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
total = 0
n = -1
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if r[ix["# Samples"]] == "# Samples":        # next kernel of a multi-kernel report: keep only the last one
        data, total, n = [], 0, -1
        continue
    n += 1
    s = int(r[ix["# Samples"]] or 0)
    total += s
    reasons = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    data.append((s, n, r[ix["Address"]], r[ix["Source"]][:70], reasons, r[ix["Instructions Executed"]], r[ix["L1 Wavefronts Shared"]]))
print("total samples", total)
# cumulative by region of 100 instructions
for lo in range(0, len(data), 100):
    print(f"instr {lo:5d}-{lo+99:5d}: {sum(d[0] for d in data[lo:lo+100]) / max(total,1) * 100:5.1f} %")
for s, n, a, src, reasons, ex, wf in sorted(data, reverse=True)[:top]:
    print(f"{s:6d} {s / total * 100:4.1f}%  #{n:5d} {src:70s} {reasons} exec={ex} wf={wf}")
