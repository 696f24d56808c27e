"""Dev tool: warp-stall picture of one `ncu --set full --import-source on` report (run where ncu is installed, no GPU needed):
stall reasons over all samples, samples by opcode, and samples per 600-instruction code region.
usage: python tools/ncu_stalls.py report.ncu-rep > profiles/xxx.md"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U, V = rows[0], rows[1], rows[2]
d = dict(zip(H, V))
print(f"# {d['Kernel Name'].split('(')[0]}\n")
for k in ('gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors_srcunit_tex.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread'):
    print(f"* {k} = {d.get(k)} {U[H.index(k)] if k in H else ''}")
for h in H:
    if h.endswith('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'):
        print(f"* {h} = {d[h]} %")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
H = rows[1]
data = [r for r in rows[2:] if len(r) == len(H)]
isrc, ins, iex = H.index('Source'), H.index('# Samples'), H.index('Instructions Executed')
stalls = [h for h in H if h.startswith('stall_') and 'Not' not in h]
tot = sum(int(r[ins] or 0) for r in data)
print(f"\nSASS instructions: {len(data)}; warp-state samples: {tot}\n\n## stall reasons (all warps)\n")
st = collections.Counter()
for r in data:
    for s in stalls:
        st[s] += int(r[H.index(s)] or 0)
for k, v in st.most_common(10):
    print(f"* {k[6:]}: {v} ({100 * v / max(sum(st.values()), 1):.1f} %)")
print("\n## samples by opcode\n")
agg, ex = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    agg[op] += int(r[ins] or 0)
    ex[op] += int(r[iex] or 0)
for op, c in agg.most_common(16):
    print(f"* {op}: {100 * c / tot:.1f} % of samples, {ex[op]} warp-instructions executed")
print("\n## samples per 600-instruction region (address order: epilogue threads first, then producer / issuer)\n")
for c0 in range(0, len(data), 600):
    ch = data[c0:c0 + 600]
    s = sum(int(r[ins] or 0) for r in ch)
    cs = collections.Counter()
    for r in ch:
        for x in stalls:
            cs[x] += int(r[H.index(x)] or 0)
    ops = collections.Counter((t[1] if t[0].startswith('@') else t[0]).split('.')[0] for t in (r[isrc].split() for r in ch) if t)
    print(f"* {c0}: {s} samples; top stalls {[(k[6:], v) for k, v in cs.most_common(3)]}; top opcodes {[k for k, _ in ops.most_common(4)]}")
