"""`ncu -i <rep> --page raw --csv` on stdin -> one line per kernel (its last launch): time, DRAM bytes, pipe utilisation, top stall reasons."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
last = {}
for r in rows[2:]:
    last[r[col['Kernel Name']]] = r


def g(r, k):
    try:
        return float(r[col[k]])
    except Exception:  # noqa: BLE001
        return float('nan')


print('kernel | time us | dram read MB | dram write MB | dram % | tensor pipe % | issue active % | warps active % | regs | top stalls (warps per issue)')
for name, r in last.items():
    st = sorted(((g(r, h), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for h in hdr
                 if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')), reverse=True)[:3]
    print(' | '.join([name.split('(')[0][:60], f"{g(r, 'gpu__time_duration.sum'):.1f}", f"{g(r, 'dram__bytes_read.sum'):.1f}", f"{g(r, 'dram__bytes_write.sum'):.1f}",
                      f"{g(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}", f"{g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f}",
                      f"{g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}", f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}",
                      f"{g(r, 'launch__registers_per_thread'):.0f}", ', '.join(f'{n} {v:.2f}' for v, n in st)]))
