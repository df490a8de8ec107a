#!/bin/bash
# One GPU call that measures every staged (default-off) variant against the shipped build -- run as
#   gpurun --timeout 240 -- 'bash scripts/staged_ab.sh'
# after building the variant libraries HERE (they travel with the snapshot):
#   make -C x265-yuuki-asuna_b200/csrc exp EXPNAME=hp EXPFLAGS=-DME_HPEL_PAIRS=1
#   make -C x265-yuuki-asuna_b200/csrc exp EXPNAME=vr EXPFLAGS=-DME_VCELL_REUSE=1
#   make -C x265-yuuki-asuna_b200/csrc exp EXPNAME=hpvr EXPFLAGS='-DME_HPEL_PAIRS=1 -DME_VCELL_REUSE=1'
#   make -C x265-yuuki-asuna_b200/csrc exp EXPNAME=la EXPSRC=la_search_thread EXPFLAGS=-DLA_PACKED_SATD=1
# Results: gpurun_out/ab_me_frame.log, ab_interp.log, staged_la.log, staged_interp_tests.log
set -u
cd "$(dirname "$0")/.."
P=x265-yuuki-asuna_b200
mkdir -p gpurun_out
# 1. frame search: half-pel candidate pairs (equal results on 144 cases + a 2160p frame, then timing)
EXPS=""; for v in hp vr hpvr; do [ -f $P/libx265b200_$v.so ] && EXPS="$EXPS --exp $P/libx265b200_$v.so"; done
[ -n "$EXPS" ] && timeout 90 python scripts/ab_me_frame.py $EXPS
# 2. interpolation: cell form vs pixel form on the bench's 173 400 blocks, then the parity tests with the cell form on
timeout 60 python scripts/ab_interp.py
X265B200_INTERP_FAST=1 timeout 120 python -m pytest tests/test_interp_intra_gpu.py tests/test_mc_gpu.py -x -q > gpurun_out/staged_interp_tests.log 2>&1
tail -3 gpurun_out/staged_interp_tests.log
# 3. lookahead: packed-word SATD variant -- parity tests, then the lookahead timing script with each library
if [ -f $P/libx265b200_la.so ]; then
  X265B200_LIB=$PWD/$P/libx265b200_la.so timeout 120 python -m pytest tests/test_lookahead_gpu.py -x -q > gpurun_out/staged_la.log 2>&1
  tail -3 gpurun_out/staged_la.log
  for lib in $P/libx265b200.so $P/libx265b200_la.so; do
    echo "== la_bench with $lib" >> gpurun_out/staged_la.log
    X265B200_LIB=$PWD/$lib timeout 120 python scripts/la_bench.py >> gpurun_out/staged_la.log 2>&1
  done
  tail -12 gpurun_out/staged_la.log
fi
# 4. intra: cell form vs persistent shared-memory form on the bench's intra stage, then the parity tests with the cell form on
timeout 120 python scripts/ab_intra.py
X265B200_INTRA_FAST=1 timeout 120 python -m pytest tests/test_interp_intra_gpu.py tests/test_fullsize_gpu.py -k intra -x -q > gpurun_out/staged_intra_tests.log 2>&1
tail -3 gpurun_out/staged_intra_tests.log
