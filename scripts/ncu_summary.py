"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the few counters we quote."""
import csv, subprocess, sys
WANT = [('gpu__time_duration.sum', 'dur'), ('launch__grid_size', 'grid'),
        ('launch__registers_per_thread', 'regs'),
        ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'dmma_pct_active'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
        ('lts__t_sector_hit_rate.pct', 'l2_hit'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ_pct'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st_long'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'st_math'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st_bar'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st_short'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st_wait'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64_pct_active'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu_pct_active'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conflicts'),
        ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct_elapsed'),
        ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'tc_smem_wavefronts_pct'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum', 'l2_to_sm_read'),
        ('l1tex__m_xbar2l1tex_read_sectors.sum.pct_of_peak_sustained_elapsed', 'l2_to_sm_pct'),
        ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu_wavefronts_pct')]
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
ki = h.index('Kernel Name')
for r in rows[2:]:
    parts = ['%-28s' % r[ki].split('(')[0][:28]]
    for name, short in WANT:
        if name in h:
            i = h.index(name)
            parts.append('%s=%s%s' % (short, r[i][:9], u[i] if short in ('dur', 'dram_rd', 'dram_wr', 'l2_to_sm_read') else ''))
    try:  # achieved DRAM GB/s = (read + write bytes) / duration, in the units ncu printed
        def val(name):
            i = h.index(name)
            x = float(r[i].replace(',', ''))
            unit = u[i].lower()
            scale = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6,
                     'usecond': 1e-6, 'ms': 1e-3, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}
            return x * scale.get(unit, 1.0)
        gbs = (val('dram__bytes_read.sum') + val('dram__bytes_write.sum')) / val('gpu__time_duration.sum') / 1e9
        parts.append('dram_GBps=%.0f' % gbs)
    except Exception:
        pass
    print('  '.join(parts))
