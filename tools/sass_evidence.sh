#!/bin/sh
# Counts of the Blackwell tensor / TMA / TMEM / mbarrier SASS mnemonics per kernel of the built library (profiles/rNN_sass_evidence.txt).
#   tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
LIB=${1:-relightableavatar_b200/libra_b200.so}
echo "cuobjdump -sass $LIB ($(git rev-parse --short HEAD 2>/dev/null)): count mnemonic, per kernel (sm_100a)"
echo "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st, UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit,"
echo "UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier ops"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function : /{name=$3}
/UTCHMMA|UTCBAR|UBLKCP|LDTM|STTM|UTCATOMSWS|UTMALDG|UTCCP/{
  n=split($0,a," ");
  for(i=1;i<=n;i++) if (a[i] ~ /^(UTCHMMA|UTCBAR|UBLKCP|LDTM|STTM|UTCATOMSWS|UTMALDG|UTCCP)/) { gsub(/[;,]/,"",a[i]); c[name" "a[i]]++ }
}
END{for(k in c) printf "%6d  %s\n", c[k], k}' | sort -k2,2 -k3,3 | c++filt 2>/dev/null
