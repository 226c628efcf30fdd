#!/bin/bash
D=gpurun_out/$1; mkdir -p $D
timeout 300 python tools/profile_all.py > $D/plain.log 2>&1 || { tail -5 $D/plain.log; exit 1; }
timeout 1500 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:hmmc:: -o $D/all -f python tools/profile_all.py > $D/ncu.log 2>&1
tail -3 $D/ncu.log; ls -la $D
