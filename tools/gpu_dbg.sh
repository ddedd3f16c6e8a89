#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /dev/shm/c
OUT=gpurun_out/dbg.log
: > $OUT
REF=oracle/_ref/kvazaar; CTU=oracle/_ref/kvazaar_ctu; LIB=$PWD/kvazaar_b200/libkvzcuda.so
python tools/synth_yuv.py 1920 1080 30 /dev/shm/c/a1080.yuv
$REF -i /dev/shm/c/a1080.yuv --input-res 1920x1080 -o /dev/shm/c/ref.hevc --preset medium -q 27 -p 1 --owf 5 2>/dev/shm/c/ref.err
echo "== replace, owf 5 (6 slots, 30 frames)" >> $OUT
KVZ_CTU_PROVIDER=$LIB timeout 300 $CTU -i /dev/shm/c/a1080.yuv --input-res 1920x1080 -o /dev/shm/c/ctu.hevc --preset medium -q 27 -p 1 --owf 5 2>/dev/shm/c/ctu.err
echo "rc=$?" >> $OUT; tail -5 /dev/shm/c/ctu.err >> $OUT
cmp /dev/shm/c/ref.hevc /dev/shm/c/ctu.hevc >> $OUT 2>&1 && echo IDENTICAL >> $OUT
echo "== verify, owf 5" >> $OUT
KVZ_CUDA_CTU_DEBUG=1 KVZ_CTU_MODE=verify KVZ_CTU_PROVIDER=$LIB timeout 300 $CTU -i /dev/shm/c/a1080.yuv --input-res 1920x1080 -o /dev/shm/c/ver.hevc --preset medium -q 27 -p 1 --owf 5 2>/dev/shm/c/ver.err
echo "rc=$?" >> $OUT; grep -E "kvz-ctu" /dev/shm/c/ver.err | head -30 >> $OUT
echo "== verify, diag launches, owf 5" >> $OUT
KVZ_CUDA_CTU_DIAG=1 KVZ_CUDA_CTU_DEBUG=1 KVZ_CTU_MODE=verify KVZ_CTU_PROVIDER=$LIB timeout 300 $CTU -i /dev/shm/c/a1080.yuv --input-res 1920x1080 -o /dev/shm/c/ver.hevc --preset medium -q 27 -p 1 --owf 5 2>/dev/shm/c/ver.err
echo "rc=$?" >> $OUT; grep -E "kvz-ctu" /dev/shm/c/ver.err | head -30 >> $OUT
cat $OUT
