#!/bin/bash
# SASS mnemonic counts per object file of libsvanon_b200.so (run after __graft_entry__.build(); no GPU needed).
# Usage: tools/sass_mnemonics.sh > profiles/<tag>_sass_mnemonics.txt
cd "$(dirname "$0")/.."
echo "SASS mnemonic counts per object file of libsvanon_b200.so (cuobjdump -sass streamvoiceanon_b200/build/*.o, sm_100a, repo at $(git rev-parse --short HEAD)):"
echo "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (tensor-map TMA; .2D / .3D = tensor rank), UBLKCP = cp.async.bulk (1-D bulk copy),"
echo "LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, R2UR.BROADCAST = per-lane descriptor hand-over"
echo
for o in streamvoiceanon_b200/build/*.o; do
  c=$(cuobjdump -sass "$o" | grep -oE "\b(UTCHMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|LDTM[.a-zA-Z0-9_]*|STTM[.a-zA-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|R2UR\.BROADCAST)" | sort | uniq -c | awk '{printf "%s%s x%s", (NR>1?", ":""), $2, $1}')
  [ -n "$c" ] && echo "$(basename "$o"): $c"
done
