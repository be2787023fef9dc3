"""Developer tool: condenses `ncu --page raw --csv` of a report into a small JSON for profiles/.
    python tools/ncu_summary.py report.ncu-rep out.json [note]"""
import csv, json, subprocess, sys
KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_on.sum', 'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum', 'smsp__sass_inst_executed_op_utcmma.sum', 'smsp__sass_inst_executed_op_utccp.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
recs = []
for d in rows[2:]:
    rec = {k: (d[hdr.index(k)] + ' ' + units[hdr.index(k)]).strip() for k in KEEP if k in hdr}
    if len(sys.argv) > 3:
        rec['note'] = sys.argv[3]
    recs.append(rec)
json.dump(recs, open(sys.argv[2], 'w'), indent=1)
for r in recs:
    print(r['Kernel Name'][:80], r.get('gpu__time_duration.sum'), 'dram rd', r.get('dram__bytes_read.sum'), 'wr', r.get('dram__bytes_write.sum'))
