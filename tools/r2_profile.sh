#!/bin/bash
# usage: r2_profile.sh <outdir>  -- ncu --set full of every shipped kernel (one launch each), exported as CSV on the box
D=gpurun_out/$1; mkdir -p $D
REP=2 timeout 300 python tools/profile_all.py > $D/plain.log 2>&1 || { tail -5 $D/plain.log; exit 1; }
REP=1 timeout 1200 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:hmmc:: -o /tmp/all -f python tools/profile_all.py > $D/ncu.log 2>&1
tail -2 $D/ncu.log; ls -la /tmp/all.ncu-rep
ncu -i /tmp/all.ncu-rep --page raw --csv > $D/all_raw.csv 2>/dev/null
ncu -i /tmp/all.ncu-rep --page details --csv > $D/all_details.csv 2>/dev/null
SZ=$(stat -c %s /tmp/all.ncu-rep); if [ $SZ -lt 40000000 ]; then cp /tmp/all.ncu-rep $D/all.ncu-rep; fi
ls -la $D
