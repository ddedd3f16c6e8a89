#!/bin/bash
# GPU box: the reference CLI with the CTU-driver hooks bound to libkvzcuda.so; verify + bitstream identity + timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/c
OUT=gpurun_out/ctu_check.log
: > $OUT
REF=oracle/_ref/kvazaar; CTU=oracle/_ref/kvazaar_ctu; LIB=$PWD/kvazaar_b200/libkvzcuda.so
nvidia-smi -L >> $OUT 2>&1
nproc >> $OUT
run() {  # name w h frames preset qp [noisy]
  local f=/tmp/c/$1.yuv
  [ -f $f ] || python tools/synth_yuv.py $2 $3 $4 $f $7
  $REF -i $f --input-res $2x$3 -o /tmp/c/ref.hevc --preset $5 -q $6 -p 1 2>/tmp/c/ref.err
  KVZ_CUDA_CTU_DEBUG=1 KVZ_CTU_PROVIDER=$LIB KVZ_CTU_MODE=verify timeout 600 $CTU -i $f --input-res $2x$3 -o /tmp/c/ver.hevc --preset $5 -q $6 -p 1 2>&1 | grep "kvz-ctu" | grep -v active | head -12 >> $OUT
  KVZ_CTU_PROVIDER=$LIB timeout 600 $CTU -i $f --input-res $2x$3 -o /tmp/c/rep.hevc --preset $5 -q $6 -p 1 2>/tmp/c/rep.err
  if cmp -s /tmp/c/ref.hevc /tmp/c/rep.hevc; then echo "IDENTICAL $* bytes=$(stat -c %s /tmp/c/ref.hevc) ref_fps=$(grep FPS /tmp/c/ref.err | awk '{print $2}') ctu_fps=$(grep FPS /tmp/c/rep.err | awk '{print $2}')" >> $OUT
  else echo "DIFFERENT $* $(tail -3 /tmp/c/rep.err | tr '\n' ' ')" >> $OUT; fi
}
run a64 64 64 3 ultrafast 32
run a264 264 200 2 medium 27
run a264 264 200 2 veryslow 22
run n264 264 200 2 veryslow 22 --noisy
run n264 264 200 2 medium 27 --noisy
run a832 832 480 4 veryslow 22
if [ "$1" != "quick" ]; then
run a1080 1920 1080 16 medium 27
run a2160 3840 2160 8 veryslow 22
fi
cat $OUT
