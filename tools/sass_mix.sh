#!/bin/sh
# SASS evidence per kernel (no GPU needed): the instructions that prove the intended hardware paths -- FP64 tensor-core
# tiles (DMMA), TMA loads (UTMALDG), 128-bit global loads, global / shared reductions, cluster barriers -- and the ten most
# frequent mnemonics.  usage: tools/sass_mix.sh > profiles/rN_sass_mix.txt
cd "$(dirname "$0")/.."
for k in k_fused_hist k_span_hist k_tile_hist k_vote_refs k_fused_fixup_stats k_boot_classify k_boot_split k_bootstrap_mmaILj0ELb1 k_bootstrap_mmaILj0ELb0 k_bootstrapILb1 k_momentsILb0 k_pack k_hist k_site_hist k_window_rstats; do
  sh tools/sass_of.sh "$k" > /tmp/sass_one.txt
  n=$(wc -l < /tmp/sass_one.txt)
  [ "$n" -eq 0 ] && continue
  echo "== $k ($n instructions in all instantiations matched)"
  for m in 'DMMA' 'UTMALDG' 'LDG\.E[A-Z.]*\.128 ' 'REDG' 'ATOMG' 'ATOMS' 'MATCH' 'SHFL' 'UCGABAR' 'DFMA|DMUL|DADD' 'MUFU'; do
    c=$(grep -cE " ($m)" /tmp/sass_one.txt)
    [ "$c" -gt 0 ] && printf "   %-28s %6d\n" "$m" "$c"
  done
  echo "   top mnemonics: $(awk '{print $2}' /tmp/sass_one.txt | sed 's/\..*//' | sort | uniq -c | sort -rn | head -10 | awk '{printf "%s x%s  ", $2, $1}')"
done
