#!/bin/bash
# One gpurun call: ncu launch list of one ADPM2 iteration, DRAM traffic of 1 and 2 iterations, ncu --set full of the dominant
# kernels, for one precision mode (default fp16, the default mode).  Everything lands in gpurun_out/ (summaries are copied to
# profiles/ by hand after reading them here).   Usage: tools/final_measure.sh [fp16|tf32|bf16] [tag]
prec=${1:-fp16}; tag=${2:-r02}
case $prec in tf32) K=1;; bf16) K=2;; *) K=3;; esac
mkdir -p gpurun_out
MDT_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_launches_$prec.csv python tools/profile_step.py $prec 4096 2 > gpurun_out/ncu_l.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_$prec.csv > gpurun_out/${tag}_launches_$prec.summary.txt 2>&1; head -14 gpurun_out/${tag}_launches_$prec.summary.txt
tools/measure_traffic.sh $prec 4096
if [ "$3" = "lists" ]; then exit 0; fi
cap() {  # name regex skip count
  MDT_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o gpurun_out/${tag}_prof_${prec}_$1 -f python tools/profile_step.py $prec 4096 2 > gpurun_out/ncu_$1.log 2>&1
  ls -la gpurun_out/${tag}_prof_${prec}_$1.ncu-rep 2>&1 | cut -c30-
}
cap attn_frag_self "attn_frag_kernel<\\(int\\)$K, \\(int\\)0," 4 2
cap attn_frag_cross "attn_frag_kernel<\\(int\\)$K, \\(int\\)16," 2 2
cap attn_frag_cross_l2 "attn_frag_kernel<\\(int\\)$K, \\(int\\)4," 2 2
cap ff_chain 'ff_chain_kernel' 2 2
cap gemm_tma "gemm_tma_kernel<\\(int\\)$K, \\(bool\\)0>" 20 8
cap gemm_tma_ln "gemm_tma_kernel<\\(int\\)$K, \\(bool\\)1>" 6 4
cap gn_apply 'gn_apply_(reg|slab)_kernel' 4 6
cap resnet_small 'resnet_small_kernel' 0 2
python tools/ncu_summary.py gpurun_out/${tag}_ncu_summary_$prec.csv gpurun_out/${tag}_prof_${prec}_*.ncu-rep
# only text and the two reports worth reading source-level travel back (gpurun_out/ is capped at 64 MiB)
for f in gpurun_out/${tag}_prof_${prec}_*.ncu-rep; do case $f in *attn_frag_self*|*gemm_tma_ln*) ;; *) rm -f $f;; esac; done
du -sh gpurun_out
