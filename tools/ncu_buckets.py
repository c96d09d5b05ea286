"""Bucket the SASS-level samples of an ncu report by instruction index: where the warps of a kernel spend their time."""
import csv, io, subprocess, sys
rep = sys.argv[1]; frames = float(sys.argv[2]); B = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[-1]
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
          'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'launch__registers_per_thread'] + [k for k in h if 'issue_stalled' in k and 'per_issue' in k]:
    if k in h:
        i = h.index(k); print(f"{k.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio','')},{u[i]},{v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ci = {k: i for i, k in enumerate(hdr)}; data = rows[2:]
tot_s = sum(int(r[ci['# Samples']] or 0) for r in data)
print("# total samples", tot_s, "warp-inst/frame", sum(int(r[ci['Instructions Executed']]) for r in data) / frames)
for a in range(0, len(data), B):
    seg = data[a:a + B]
    s = sum(int(r[ci['# Samples']] or 0) for r in seg); n = sum(int(r[ci['Instructions Executed']]) for r in seg)
    if s < 0.004 * tot_s: continue
    ops = [(r[ci['Source']].split()[1] if r[ci['Source']].strip().startswith('@') else r[ci['Source']].split()[0]) for r in seg]
    key = sorted({o for o in ops if any(t in o for t in ('UTC', 'SYNCS', 'BAR', 'LDTM', 'UBLKCP', 'MUFU', 'STG', 'ATOM', 'F2F', 'LDS.128', 'STS.128'))})
    top = max(seg, key=lambda r: int(r[ci['# Samples']] or 0))
    print(f"[{a:4d}] {100*s/tot_s:5.1f}% inst/frame {n/frames:6.1f} {key} | hottest: {top[ci['Source']].strip()[:60]} ({top[ci['# Samples']]})")
