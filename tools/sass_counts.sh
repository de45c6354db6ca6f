#!/bin/bash
# SASS mnemonic counts + ptxas resource maxima of the built objects (no GPU needed): the evidence that the hot ops are tcgen05 / TMA-bulk
# code (B200_PROFILING.md: UTCHMMA / UTCBAR / LDTM / STTM / UBLKCP).  usage: tools/sass_counts.sh > profiles/rNN_sass_mnemonics.txt
cd "$(dirname "$0")/../nbss_b200/csrc" || exit 1
make -s -j >/dev/null || exit 1
echo "# SASS evidence of the build at $(git rev-parse --short HEAD) (cuobjdump -sass nbss_b200/csrc/build/*.o, sm_100a; tools/sass_counts.sh):"
echo "# tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, tcgen05.ld = LDTM, tcgen05.st = STTM, tcgen05.alloc/dealloc = UTCATOMSWS,"
echo "# cp.async.bulk (TMA bulk copy) = UBLKCP, cp.async.bulk.prefetch.L2 = UBLKPF, elect.sync = ELECT, mbarrier ops = SYNCS, MUFU = transcendentals,"
echo "# R2UR = register -> uniform-register moves (descriptor staging; 429 in ffn_fwd before the warp-uniform issue fix);"
echo "# max_regs / max_spill_store_bytes: maximum over the kernels of the object, from -Xptxas -v (build/*.ptxas.log)"
for f in ffn_fwd ffn_bwd mhsa_fwd mhsa_bwd wgrad fconv_tc fullband_tc fullband_rows_tc online io loss predict crossband pack; do
    cuobjdump -sass build/$f.o > /tmp/_$f.sass
    line="$f"
    for m in UTCHMMA UTCBAR LDTM STTM UBLKCP UBLKPF UTCATOMSWS ELECT SYNCS MUFU R2UR; do
        c=$(grep -c "^\s*/\*[0-9a-f]*\*/\s*\(@!\?U\?P[0-9T]\s\+\)\?$m" /tmp/_$f.sass)
        line="$line $m=$c"
    done
    regs=$(grep -o "Used [0-9]* registers" build/$f.ptxas.log | awk '{print $2}' | sort -n | tail -1)
    sp=$(grep -o "[0-9]* bytes spill stores" build/$f.ptxas.log | awk '{print $1}' | sort -n | tail -1)
    echo "$line max_regs=$regs max_spill_store_bytes=$sp"
    rm -f /tmp/_$f.sass
done
