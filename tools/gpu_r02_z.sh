#!/bin/bash
# round 2: A/B of three builds of the library on one box, alternating processes:
#   B = in-tree (cta_group::2 pairs + packed-fp32 epilogue), A = the previous commit's tc_gemm.cu, C = in-tree with scalar epilogue math
#   (the two extra libraries are built by hand into ab_libs/ — see DESIGN.md 4b; they are not kept in the tree)
mkdir -p gpurun_out; : > gpurun_out/r02_z_libs.jsonl
for rep in 1 2; do
  for lib in "" ab_libs/libA_old.so ab_libs/libC_scalar.so; do
    AB_LIB=$lib AB_ROUNDS=2 timeout 120 python tools/step_ab.py default= single=MASR_TC_PAIR:0 >> gpurun_out/r02_z_libs.jsonl 2>> gpurun_out/r02_z_libs.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/r02_z_libs.jsonl"):
    d = json.loads(l); print(d["lib"], d["identical_outputs"], {k: v["median"] for k, v in d["ms_per_step"].items()})
PY
