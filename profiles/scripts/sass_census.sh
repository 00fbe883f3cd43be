#!/bin/bash
# SASS census of liballrank_b200.so per kernel: tcgen05 MMAs (UTC*MMA), TMEM loads/stores (LDTM/STTM), TMA
# (UTMALDG/UTMASTG), mbarrier waits (SYNCS), legacy tensor instructions (HMMA) -- evidence that the hot kernels are
# tcgen05/TMEM/TMA code (B200_PROFILING.md "What proves a Blackwell-native kernel").  CPU-only (cuobjdump).
SO=${1:-allrank_b200/liballrank_b200.so}
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; sub(/^_Z[0-9]*/, "", fn); name[fn]=1; cur=fn; next }
  cur != "" {
    if ($0 ~ /UTC[A-Z]*MMA/) mma[cur]++
    if ($0 ~ /LDTM/) ldtm[cur]++
    if ($0 ~ /STTM/) sttm[cur]++
    if ($0 ~ /UTMALDG/) tmald[cur]++
    if ($0 ~ /UTMASTG/) tmast[cur]++
    if ($0 ~ /SYNCS/) syncs[cur]++
    if ($0 ~ /HMMA/ && $0 !~ /UTC/) hmma[cur]++
    if ($0 ~ /MUFU/) mufu[cur]++
    if ($0 ~ /RED\.|REDG|ATOMG|ATOMS|REDUX/) red[cur]++
  }
  END {
    printf "%-110s %6s %5s %5s %7s %7s %6s %5s %5s %5s\n", "kernel (mangled, truncated)", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MUFU", "RED"
    for (k in name) if (mma[k] + ldtm[k] + tmald[k] + tmast[k] + mufu[k] + red[k] > 0)
      printf "%-110s %6d %5d %5d %7d %7d %6d %5d %5d %5d\n", substr(k, 1, 110), mma[k], ldtm[k], sttm[k], tmald[k], tmast[k], syncs[k], hmma[k], mufu[k], red[k]
  }' | (read -r hdr; echo "$hdr"; sort)
