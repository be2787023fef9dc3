"""Developer tool: per-kernel SASS opcode counts of the built library -> profiles/sass_opcodes.txt
    python tools/sass_opcodes.py"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "libxsmm-1_b200", "lib", "libxsmm_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ['UTCHMMA', 'UTCQMMA', 'UTCCP', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'SYNCS', 'FHFMA', 'FFMA', 'DFMA', 'HMMA', 'LDGSTS', 'REDUX', 'UTCBAR']
rows = []
for p in re.split(r'\n\s*Function : ', txt)[1:]:
    name = p.split('\n', 1)[0].strip()
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    dem = re.sub(r'\(CUtensorMap_st.*|\(xb::.*|\(void.*', '', dem)[:112]
    ops, n = collections.Counter(), 0
    for line in p.split('\n'):
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            n += 1
            op = m.group(1)
            for k in keys:
                if op == k or op.startswith(k + '.'):
                    ops[k] += 1
            if op.startswith('UTCHMMA') and '2CTA' in op:
                ops['UTCHMMA.2CTA'] += 1
    rows.append((dem, n, ops))
out = ["# SASS opcode counts per kernel of libxsmm-1_b200/lib/libxsmm_b200.so (cuobjdump -sass, sm_100a)",
       "# tcgen05.mma.sp (2:4 structured-sparse A, spmdm_compute_tc16s_kernel) is the SAME opcode: sparsity is bit 2 of the instruction descriptor; what tells it apart",
       "# is UTCCP = tcgen05.cp (the metadata image -> tensor memory) next to it and ncu's sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_on counter",
       "# (profiles/r02_ncu_full_c2_compute_tc16s.json).  UTMASTG = TMA tensor store,",
       "# UTCHMMA = tcgen05.mma (UTCHMMA.2CTA: cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load (cp.async.bulk.tensor),",
       "# UTMAPF = TMA L2 prefetch, SYNCS = mbarrier operations, FHFMA = mixed-precision fma.rn.f32.bf16 (K2s), DFMA = fp64 fma, LDGSTS = cp.async.",
       "# The baked fsspmdm kernels (register and strip form) are emitted as PTX at create time and assembled by the driver: they are not",
       "# in the library image (source of a given operator's kernel: libxsmm_b200_fsspmdm_kernel_source; the strip form uses UTMALDG + SYNCS).",
       "%-114s %7s  %s" % ("kernel", "instrs", "opcodes of interest")]
for dem, n, ops in sorted(rows):
    out.append("%-114s %7d  %s" % (dem, n, ' '.join('%s=%d' % (k, v) for k, v in sorted(ops.items()))))
open(os.path.join(ROOT, "profiles", "sass_opcodes.txt"), "w").write('\n'.join(out) + '\n')
print("\n".join(out[5:]))
