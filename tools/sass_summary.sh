#!/bin/bash
# Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (profiles/sass_summary.txt):
#   UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA engine),
#   UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FADD2 / FFMA2 = packed fp32x2 arithmetic, MUFU = sin / cos / ex2.
cd "$(dirname "$0")/.."
lib=mri-super-resolution_b200/lib/libb200inr.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  $lib  ($(git rev-parse --short HEAD 2>/dev/null))"
printf "%-46s %8s %6s %6s %6s %7s %7s %6s %6s %6s\n" kernel UTCHMMA .2CTA LDTM STTM UBLKCP UTCBAR FADD2 FFMA2 MUFU
cuobjdump -sass "$lib" 2>/dev/null | c++filt | awk '
  /Function : / { if (name != "") flush(); name = $0; sub(/.*Function : /, "", name); sub(/\(.*/, "", name);
                  gsub(/b200inr::/, "", name); sub(/^void /, "", name); for (k in c) delete c[k] }
  /UTCHMMA/ { c["m"]++; if ($0 ~ /2CTA/) c["p"]++ }
  /LDTM/ { c["l"]++ } /STTM/ { c["s"]++ } /UBLKCP/ { c["b"]++ } /UTCBAR/ { c["c"]++ }
  /FADD2/ { c["a2"]++ } /FFMA2/ { c["f2"]++ } /MUFU/ { c["u"]++ }
  function flush() { printf "%-46s %8d %6d %6d %6d %7d %7d %6d %6d %6d\n", substr(name, 1, 46), c["m"], c["p"], c["l"], c["s"], c["b"], c["c"], c["a2"], c["f2"], c["u"] }
  END { if (name != "") flush() }' | sort
