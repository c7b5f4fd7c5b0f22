#!/bin/bash
# output_all at 1e8 markers: device time of the limb and the CAS histogram kernels, per-launch ncu durations, and one
# `ncu --set full` capture of the limb kernel at 2e7 markers (summarised on the box)
python tools_py3/dev/time_diag.py 1e8
PIC1DP_DIAG_CAS=1 python tools_py3/dev/time_diag.py 1e8
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_diag|k_absmax" -c 12 --csv --log-file gpurun_out/r02u_diag_launches.csv python tools_py3/dev/time_diag.py 1e8 > gpurun_out/r02u_diag_ncu.log 2>&1
grep -E "k_diag|k_absmax" gpurun_out/r02u_diag_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
ncu --set full --clock-control none --import-source on -k regex:"k_diag_limb$" -s 1 -c 1 -f -o gpurun_out/r02u_prof_diag_limb python tools_py3/dev/time_diag.py 2e7 > gpurun_out/r02u_ncu_full.log 2>&1
python tools_py3/ncu_summary.py gpurun_out/r02u_prof_diag_limb.ncu-rep > gpurun_out/r02u_ncu_diag_limb.md
rm -f gpurun_out/r02u_prof_diag_limb.ncu-rep
