"""Dev tool: turn an ncu launch-list CSV / a --set full report into the markdown summaries kept under profiles/."""
import collections, csv, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_pipe_uniform.sum']


def launches(path, title, cmd, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        v = float(r[vi].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(r[ui], 1)
        k = r[ki].split('(')[0][:90]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    o = [f"# {title}", f"command: {cmd}",
         f"launches: {len(data)}   total device time: {tot / 1e6:.1f} ms (cold-cache, serialised: compare shares, not absolutes)", "",
         "| share | total ms | launches | kernel |", "|---|---|---|---|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
        o.append(f"| {100 * v / tot:.3f}% | {v / 1e6:.3f} | {n} | `{k}` |")
    open(out, 'w').write('\n'.join(o) + '\n')


def full(rep, title, cmd, notes, out):
    raw = open(rep).read() if rep.endswith('.csv') else \
        subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U, V = rows[0], rows[1], rows[2]
    o = [f"# {title}", f"command: {cmd}", "", "| metric | unit | value |", "|---|---|---|"]
    for i, h in enumerate(H):
        hh = h.split('.', 2)[-1] if h.split('.')[0] in ('LTS', 'SM_A', 'SM_B', 'SM_C', 'TPC') else h
        if (h in KEEP or hh in KEEP) and V[i] != '':
            o.append(f"| {h} | {U[i]} | {V[i]} |")
    o += [""] + notes
    open(out, 'w').write('\n'.join(o) + '\n')


if __name__ == '__main__':
    # round 1, final (f16f8 default): produced by tools/gpu_evidence.sh
    launches('gpurun_out/r01e_launches_bench.csv', 'ncu launch list, round 1 final (f16f8 operand scheme, CTA-pair tcgen05 engine)',
             'ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline',
             'profiles/r01e_launches_bench.md')
    full('gpurun_out/r01e_image_f16f8_raw.csv', 'ncu --set full, image_umma_kernel<PAIR=1, SCHEME=f16f8>, round 1 final',
         'ncu --set full --clock-control none --import-source on -k regex:image_umma -s 1 -c 1 python bench.py --batch 8 --res 1024 --steps 1 --warmup 1 --no-cpu-baseline',
         ['workload of the captured launch: batch 8 @ 1024x1024 = 8,388,608 coords = 65,536 tiles of 128, 74 CTA pairs (148 CTAs)',
          'algorithmic DRAM bytes: planes once (176.2 MB) + 12 B/coord out (100.7 MB) = 276.8 MB'],
         'profiles/r01e_image_umma_full.md')
    for w, note in (('occupancy', 'batch 32 x (128^3 grid + 100k random points) = 70,308,864 points'),
                    ('video', 'batch 16 x 256x256x16 = 16,777,216 voxels'),
                    ('nerf', 'batch 16 objects x 128x128 rays x 128 samples = 33,554,432 samples, compositing fused')):
        full(f'gpurun_out/r01e_{w}_f16f8_raw.csv', f'ncu --set full, {w}_umma_kernel (f16f8), round 1 final',
             f'ncu --set full --clock-control none -k regex:{w}_umma -s 1 -c 1 python bench.py --workload {w} --steps 1 --warmup 1',
             ['workload of the captured launch: ' + note], f'profiles/r01e_{w}_umma_full.md')
