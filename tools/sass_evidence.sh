#!/bin/bash
# SASS evidence of the Blackwell-native paths: mnemonic counts per object file and a short excerpt around the first MMA.
# usage: bash tools/sass_evidence.sh > profiles/r02_sass_evidence.txt     (after `make -C knode-cosserat_b200/csrc`)
cd "$(dirname "$0")/../knode-cosserat_b200/csrc"
for o in kc_knode_tc.o kc_train_tc3.o kc_train_tc2.o kc_train_tc.o kc_tc.o; do
  echo "==== $o ($(date -r $o +%F' '%T))"
  cuobjdump -sass $o > /tmp/_sass.txt
  for m in UTCHMMA UTCBAR LDTM STTM UBLKCP SYNCS.ARRIVE SYNCS.PHASECHK ELECT R2UR FFMA2 FMUL2 FADD2 MUFU.EX2 F2FP; do
    printf "  %-16s %6d\n" $m $(grep -c "$m" /tmp/_sass.txt)
  done
  echo "  kernels:"; grep "Function :" /tmp/_sass.txt | sed 's/^\s*/    /'
done
echo "==== excerpt: kc_knode_tc_fwd_kernel<true>, issue of the forward GEMM2 (A operand in TMEM: 'tmem[URx]' first operand)"
F=$(cuobjdump -sass kc_knode_tc.o | grep "Function : _Z22kc_knode_tc_fwd_kernelILb1" | sed 's/.*Function : //')
cuobjdump -sass -fun "$F" kc_knode_tc.o | grep -E "UTCHMMA|UTCBAR|LDTM|STTM" | sed 's#/\* 0x[0-9a-f]* \*/##' | awk '{$1=$1};1' | head -24
echo "==== excerpt: kc_train_tc3_kernel, batched issue of a 3-pass gradient product (A operand in TMEM, one ELECT per 12 UTCHMMA)"
F=$(cuobjdump -sass kc_train_tc3.o | grep "Function : _Z19kc_train_tc3_kernelILb0" | head -1 | sed "s/.*Function : //")
cuobjdump -sass -fun "$F" kc_train_tc3.o | grep -E "UTCHMMA|UTCBAR|LDTM|STTM|ELECT|UBLKCP" | sed 's#/\* 0x[0-9a-f]* \*/##' | awk '{$1=$1};1' | awk '/UTCHMMA/{f=1} f' | sed -n 1,44p
echo "==== ptxas -v (registers / spills) of the tensor-core march kernels"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr --extended-lambda -Xptxas -v -c kc_knode_tc.cu -o /tmp/_k.o 2>&1 | grep -E "Compiling entry|Used|spill" | grep -A2 "kc_knode_tc_[fb]wd" | sed 's/ptxas info    : //'
echo "==== ptxas -v of the training kernel"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr --extended-lambda -Xptxas -v -c kc_train_tc3.cu -o /tmp/_k3.o 2>&1 | grep -E "Compiling entry|Used|spill" | grep -A2 "kc_train_tc3_kernel" | sed 's/ptxas info    : //'
