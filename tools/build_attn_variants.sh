#!/bin/bash
# Builds libkvq variants that differ only in the GQA attention kernel's launch shape (A/B runs on the GPU box):
#   tools/build_attn_variants.sh "256 2 2" "128 3 2" ...      (threads, CTAs per SM, ring stages)
# Output: efficient-llm-inference_b200/csrc/_build/variants/libkvq_t<T>_c<C>_s<S>.so ; select with KVQ_LIB=<path>.
set -e
cd "$(dirname "$0")/../efficient-llm-inference_b200/csrc"
make -j8 >/dev/null
mkdir -p _build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -I../../include -I."
for v in "$@"; do
  set -- $v
  name="t$1_c$2_s$3"
  ( nvcc $FLAGS -DKVQ_ATTN_GQA_THREADS=$1 -DKVQ_ATTN_GQA_CTAS_PER_SM=$2 -DKVQ_ATTN_GQA_STAGES=$3 ${EXTRA} -c attn.cu -o _build/variants/attn_$name.o &&
    nvcc -shared -cudart static -o _build/variants/libkvq_$name.so _build/common.o _build/quant.o _build/evict.o _build/pool.o _build/variants/attn_$name.o &&
    echo built $name ) &
done
wait
