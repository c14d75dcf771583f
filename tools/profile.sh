#!/bin/bash
# One gpurun call's worth of evidence for a round (dev tool): GPU tests, the default bench line, the ncu launch list of
# the same command and one full ncu capture of the dominant kernel -- each ncu pass only after its command has exited 0
# without ncu.  Run ON the GPU box, from the repo root, e.g.
#   gpurun --timeout 900 -- 'bash tools/profile.sh r2a'
# Results land in gpurun_out/<tag>_*; copy what should be judged into profiles/.
#   SKIP_TESTS=1      leave the test-suite out          NCU_FULL=0   leave the full capture out
#   BENCH_ARGS="..."  extra bench.py arguments (e.g. "--workload cfg3", "--classes 19", "--label-format ids")
set -u
tag=${1:-run}
out=gpurun_out
mkdir -p $out
short="--steps 32 --warmup 16 --repeats 2 --no-cpu-baseline --no-e2e --no-render ${BENCH_ARGS:-}"
if [ -z "${SKIP_TESTS:-}" ]; then
  python -m pytest tests -m gpu -x -q -n 3 2>&1 | tail -15 > $out/${tag}_gpu_tests.log
  tail -3 $out/${tag}_gpu_tests.log
fi
python bench.py ${BENCH_ARGS:-} > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
cut -c1-400 $out/${tag}_bench.json
python bench.py $short > /dev/null 2> $out/${tag}_short.err || { tail -5 $out/${tag}_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py $short > $out/${tag}_ncu_list.log 2>&1
if [ "${NCU_FULL:-1}" != "0" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_fuse -s 40 -c 2 -o $out/${tag}_k_fuse \
      python bench.py $short > $out/${tag}_ncu_full.log 2>&1
  # dram bytes of the captured launches (roofline.traffic): read here, or later with `ncu -i ... --page raw --csv`
  ncu -i $out/${tag}_k_fuse.ncu-rep --page raw --csv 2>/dev/null \
    | python -c 'import csv,sys; r=list(csv.reader(sys.stdin)); h=r[0]; i=[k for k,c in enumerate(h) if c in ("dram__bytes_read.sum","dram__bytes_write.sum","gpu__time_duration.sum","smsp__inst_executed.sum")]; [print([(h[k],row[k]) for k in i]) for row in r[2:]]' \
    > $out/${tag}_k_fuse_traffic.txt 2>/dev/null
  cat $out/${tag}_k_fuse_traffic.txt
fi
