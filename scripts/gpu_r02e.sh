#!/bin/bash
for v in "TAG=default" "TAG=nosplit SDFG_TC_SPLIT=0" "TAG=nocollapse SDFG_TC_COLLAPSE=0" "TAG=oldfwd SDFG_TC_FWD=old" "TAG=pp0 SDFG_TC_PP=0" "TAG=cg1 SDFG_TC_CG=1"; do
  env $v timeout 300 python scripts/dbg_fullsize.py 1e-4 0 2>&1 | tail -1
done
env TAG=default_t1 timeout 300 python scripts/dbg_fullsize.py 1.0 1 2>&1 | tail -1
