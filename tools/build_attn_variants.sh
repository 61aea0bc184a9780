#!/bin/bash
# Builds libkvq variants that differ only in attention-kernel compile-time knobs (A/B runs on the GPU box):
#   tools/build_attn_variants.sh name1 "-DKVQ_ATTN_MMA_CTAS_PER_SM=1 -DKVQ_ATTN_MMA_MAX_STAGES=3" name2 "-DKVQ_ATTN_MHA_STAGES=2" ...
# (the knobs are compile-time macros of attn.cu / attn_mma.cu; the library reads nothing from the environment)
# Output: efficient-llm-inference_b200/csrc/_build/variants/libkvq_<name>.so ; select with KVQ_LIB=<path>.
set -e
cd "$(dirname "$0")/../efficient-llm-inference_b200/csrc"
make -j8 >/dev/null
mkdir -p _build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -I../../include -I."
while [ $# -ge 2 ]; do
  name="$1"; defs="$2"; shift 2
  ( nvcc $FLAGS $defs -c attn.cu -o _build/variants/attn_$name.o && nvcc $FLAGS $defs -c attn_mma.cu -o _build/variants/attn_mma_$name.o &&
    nvcc -shared -cudart static -o _build/variants/libkvq_$name.so _build/common.o _build/quant.o _build/evict.o _build/pool.o _build/xchg.o \
         _build/variants/attn_$name.o _build/variants/attn_mma_$name.o 2>/dev/null &&
    rm -f _build/variants/attn_$name.o _build/variants/attn_mma_$name.o && echo built $name ) &
done
wait
