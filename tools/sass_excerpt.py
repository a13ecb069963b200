"""Write profiles/sass_excerpt_r02.txt: the SASS evidence of the sm_100a code paths, cut from the built library with
cuobjdump (DMMA main loop of the variance GEMM with its conflict-free LDS.128 fragment loads, the TMA / mbarrier / setmaxnreg
sites of the producer-consumer pipeline, the bulk-copy staging of K1 and of the latency kernel, the DMMA loop of the fused
small-model kernel)."""
import os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
LIB = os.path.join(ROOT, 'gpyopt_b200', 'lib', 'libgpo_b200.so')
OUT = os.path.join(ROOT, 'profiles', 'sass_excerpt_r02.txt')


def sass(fun):
    txt = subprocess.run(['cuobjdump', '-sass', '-fun', fun, LIB], capture_output=True, text=True).stdout
    return [re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l).rstrip() for l in txt.splitlines() if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l)]


def window(lines, pred, before, after, maxhits=1):
    out, hits = [], 0
    for i, l in enumerate(lines):
        if pred(l):
            out += lines[max(0, i - before):i + after] + ['        ...']
            hits += 1
            if hits >= maxhits:
                break
    return out


def counts(lines, ops):
    return ', '.join('%s x%d' % (o, sum(1 for l in lines if re.search(r'\b' + o, l))) for o in ops)


with open(OUT, 'w') as fh:
    def emit(title, body):
        fh.write('=' * 110 + '\n' + title + '\n' + '=' * 110 + '\n' + '\n'.join(body) + '\n\n')

    g = sass('_ZN3gpo15dgemm_nt_kernelILi2ELi64ELi4EEEv14CUtensorMap_stS1_NS_10GemmParamsE')
    emit('gpo::dgemm_nt_kernel<EPI_GRAD, BN=64, STAGES=4>  (128 x 64 tile, two CTAs per SM)  -- instruction counts',
         [counts(g, ['DMMA', 'UTMALDG', 'UTMACCTL', 'SYNCS', 'USETMAXREG', 'LDS.128', 'UBLKPF', 'BAR.SYNC', 'SHFL', 'LDG', 'STG', 'LDL', 'STL'])])
    emit('... register reallocation between the producer and the consumer warpgroup (setmaxnreg.dec 40 / inc 216)',
         [l for l in g if 'USETMAXREG' in l])
    emit('... TMA producer: tensor-map prefetch, mbarrier arrive.expect_tx, two 3-D tiled loads per stage (A 128x16, B 64x16, SWIZZLE_128B)',
         window(g, lambda l: 'UTMALDG' in l, 14, 6, 1))
    emit('... consumer main loop: mbarrier try_wait on the stage, 16-byte fragment loads, DMMA.8x8x4 (first 60 lines from the first LDS.128)',
         window(g, lambda l: 'LDS.128' in l, 6, 60, 1))
    emit('... stage release (mbarrier arrive on the empty barrier) after the last DMMA of a stage',
         window(g, lambda l: 'SYNCS.ARRIVE' in l and 'TRANS64.A1T0' in l, 3, 2, 2))
    g1 = sass('_ZN3gpo15dgemm_nt_kernelILi2ELi128ELi4EEEv14CUtensorMap_stS1_NS_10GemmParamsE')
    emit('gpo::dgemm_nt_kernel<EPI_GRAD, BN=128, STAGES=4>  (128 x 128 tile, one CTA per SM) -- instruction counts',
         [counts(g1, ['DMMA', 'UTMALDG', 'SYNCS', 'USETMAXREG', 'LDS.128', 'LDL', 'STL'])] + [l for l in g1 if 'USETMAXREG' in l])
    k1 = sass('_ZN3gpo15cov_mean_kernelILi6ELb1EEEvNS_9CovParamsE')
    emit('gpo::cov_mean_kernel<6, true> (K1): bulk-copy staging of the training tiles (cp.async.bulk -> UBLKCP) + mbarrier',
         [counts(k1, ['UBLKCP', 'SYNCS', 'DFMA', 'DMUL', 'MUFU', 'STG', 'LDS'])] + window(k1, lambda l: 'UBLKCP' in l, 8, 4, 1))
    sm = sass('_ZN3gpo16smalln_wk_kernelILi8ELb0ELi32EEEvNS_12SmallNParamsE')
    emit('gpo::smalln_wk_kernel<8, false, 32> (latency path): shared-memory ring over W^-1 filled by 16-byte cp.async (LDGSTS) copies',
         [counts(sm, ['LDGSTS', 'LDGDEPBAR', 'DEPBAR', 'DFMA', 'LDS.128', 'SHFL'])] + window(sm, lambda l: 'LDGSTS' in l, 6, 4, 1))
    fu = sass('_ZN3gpo24fused_small_value_kernelILi7ELi4EEEvNS_16FusedSmallParamsE')
    emit('gpo::fused_small_value_kernel<MF=7, NF=4> (n <= 56): covariance evaluated in the B-fragment position, DMMA against the resident L^-1',
         [counts(fu, ['DMMA', 'DFMA', 'DMUL', 'MUFU', 'LDS', 'UMOV', 'SHFL', 'BSSY'])] + window(fu, lambda l: 'DMMA' in l, 10, 24, 1))
    lb = sass('_ZN3gpo19lbfgs_update_kernelENS_11LbfgsParamsE')
    emit('gpo::lbfgs_update_kernel (device-side lock-step L-BFGS state machine) -- instruction counts',
         [counts(lb, ['DFMA', 'DMUL', 'DADD', 'MUFU', 'ATOMG', 'RED', 'BRA', 'LDG', 'STG'])])
print(OUT, os.path.getsize(OUT), 'bytes')
