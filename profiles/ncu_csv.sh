#!/bin/bash
# usage (on the GPU box, via gpurun): profiles/ncu_csv.sh <tag> <kernel-regex> <cases...>
# Runs the cases plainly first, then ONE launch per case under `ncu --set full`, converts the report
# to a raw CSV on the box and deletes the (large) .ncu-rep so that gpurun_out/ stays small.
set -e
tag=$1; regex=$2; shift 2
python profiles/prof_cases.py "$@" > gpurun_out/${tag}_plain.log 2>&1
ncu --set full --clock-control none -k "regex:${regex}" -o gpurun_out/${tag} python profiles/prof_cases.py --once "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}.ncu-rep
cat gpurun_out/${tag}_plain.log
