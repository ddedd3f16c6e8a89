#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /dev/shm/c
OUT=gpurun_out/dbg.log
: > $OUT
REF=oracle/_ref/kvazaar; CTU=oracle/_ref/kvazaar_ctu; LIB=$PWD/kvazaar_b200/libkvzcuda.so
python tools/synth_yuv.py 264 200 2 /dev/shm/c/a264.yuv
python tools/synth_yuv.py 264 200 2 /dev/shm/c/n264.yuv --noisy
python tools/synth_yuv.py 1920 1080 30 /dev/shm/c/a1080.yuv
small() { # clip preset qp
  $REF -i /dev/shm/c/$1.yuv --input-res 264x200 -o /dev/shm/c/r.hevc --preset $2 -q $3 -p 1 2>/dev/null
  KVZ_CUDA_CTU_DEBUG=1 KVZ_CTU_MODE=verify KVZ_CTU_PROVIDER=$LIB timeout 120 $CTU -i /dev/shm/c/$1.yuv --input-res 264x200 -o /dev/shm/c/v.hevc --preset $2 -q $3 -p 1 2>&1 | grep "kvz-ctu" | grep -v active | head -8 >> $OUT
  KVZ_CTU_PROVIDER=$LIB timeout 120 $CTU -i /dev/shm/c/$1.yuv --input-res 264x200 -o /dev/shm/c/c.hevc --preset $2 -q $3 -p 1 2>/dev/null
  cmp -s /dev/shm/c/r.hevc /dev/shm/c/c.hevc && echo "IDENTICAL $*" >> $OUT || echo "DIFFERENT $*" >> $OUT
}
small a264 medium 27; small a264 veryslow 22; small n264 veryslow 22; small n264 medium 27; small a264 slow 32
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
for args in "--res 1920x1080 --preset medium --frames 8 --slots 1" "--res 1920x1080 --preset medium --frames 64 --slots 32" "--res 3840x2160 --preset veryslow --frames 2 --slots 1" "--res 3840x2160 --preset veryslow --frames 16 --slots 16"; do
  timeout 600 python tools/ctu_devbench.py $args >> $OUT 2>&1
done
cat $OUT
