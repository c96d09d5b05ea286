"""Summarise an ncu report: key raw metrics + per-region (by executed-count regime) instruction/stall breakdown."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; frames = float(sys.argv[2]) if len(sys.argv) > 2 else 480256.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[-1]
keep = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.avg','sm__cycles_elapsed.avg.per_second','launch__shared_mem_per_block_dynamic','lts__t_bytes.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for i, k in enumerate(h):
    if k in keep: print(f"{k},{u[i]},{v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ends = [i for i, r in enumerate(rows) if i > 1 and r and r[0] == 'Kernel Name']      # several captured launches: keep the first
data = rows[2:ends[0]] if ends else rows[2:]
ci = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[ci['Instructions Executed']]) for r in data)
print(f"# warp-inst total {tot}  per frame {tot/frames:.1f}")
byop = collections.Counter()
for r in data:
    sass = r[ci['Source']].strip(); n = int(r[ci['Instructions Executed']])
    op = (sass.split()[1] if sass.startswith('@') else sass.split()[0]).split('.')[0]
    byop[op] += n
print("# opcode mix per frame: " + ", ".join(f"{op} {n/frames:.1f}" for op, n in byop.most_common(14)))
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
# regions: contiguous runs with (nearly) equal executed counts
seg = [int(r[ci['Instructions Executed']]) for r in data]
start = 0
print("# regions: [first,last) n_inst exec_count inst/frame samples smem_wf/frame top-stalls")
regs = []
for i in range(1, len(seg) + 1):
    if i == len(seg) or abs(seg[i] - seg[start]) > 0.02 * max(seg[start], 1):
        regs.append((start, i)); start = i
# merge small regions into buckets keyed by exec count
for (a, b) in regs:
    if (b - a) * seg[a] / frames < 2.0: continue
    smp = sum(int(data[j][ci['# Samples']] or 0) for j in range(a, b))
    wf = sum(int(data[j][ci['L1 Wavefronts Shared']] or 0) for j in range(a, b))
    st = collections.Counter()
    for j in range(a, b):
        for s in stalls: st[s] += int(data[j][ci[s]] or 0)
    print(f"[{a},{b}) {b-a} {seg[a]} {(b-a)*seg[a]/frames:.1f} {smp} {wf/frames:.1f} " + " ".join(f"{k[6:]}={n}" for k, n in st.most_common(4)))
print("# total samples", sum(int(r[ci['# Samples']] or 0) for r in data))
