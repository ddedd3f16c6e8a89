#!/bin/bash
# GPU box: identity of the CTU driver (quick set) + device-only throughput and phase profile.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /dev/shm/c
OUT=gpurun_out/step.log
: > $OUT
REF=oracle/_ref/kvazaar; CTU=oracle/_ref/kvazaar_ctu; LIB=$PWD/kvazaar_b200/libkvzcuda.so
python tools/synth_yuv.py 264 200 2 /dev/shm/c/a264.yuv
python tools/synth_yuv.py 264 200 2 /dev/shm/c/n264.yuv --noisy
small() { # clip preset qp
  $REF -i /dev/shm/c/$1.yuv --input-res 264x200 -o /dev/shm/c/r.hevc --preset $2 -q $3 -p 1 2>/dev/null
  KVZ_CUDA_CTU_DEBUG=1 KVZ_CTU_MODE=verify KVZ_CTU_PROVIDER=$LIB timeout 120 $CTU -i /dev/shm/c/$1.yuv --input-res 264x200 -o /dev/shm/c/v.hevc --preset $2 -q $3 -p 1 2>&1 | grep "kvz-ctu" | grep -v active | head -6 >> $OUT
}
small a264 medium 27; small a264 veryslow 22; small n264 veryslow 22; small n264 medium 27
python tools/synth_yuv.py 1920 1080 48 /dev/shm/c/a1080.yuv
python tools/synth_yuv.py 3840 2160 16 /dev/shm/c/a2160.yuv
big() { # clip w h preset qp frames owf
  $REF -i /dev/shm/c/$1.yuv --input-res $2x$3 -o /dev/shm/c/r.hevc --preset $4 -q $5 -p 1 -n $6 2>/dev/shm/c/r.err
  KVZ_CTU_PROVIDER=$LIB timeout 200 $CTU -i /dev/shm/c/$1.yuv --input-res $2x$3 -o /dev/shm/c/c.hevc --preset $4 -q $5 -p 1 -n $6 --owf $7 2>/dev/shm/c/c.err
  echo "rc=$? $(cmp -s /dev/shm/c/r.hevc /dev/shm/c/c.hevc && echo IDENTICAL || echo DIFFERENT) $* ref: $(grep FPS /dev/shm/c/r.err) ctu: $(grep FPS /dev/shm/c/c.err) $(grep -i "assert\|abort" /dev/shm/c/c.err | head -2)" >> $OUT
}
big a1080 1920 1080 medium 27 48 47
big a2160 3840 2160 veryslow 22 16 15
for args in "--res 1920x1080 --preset medium --frames 8 --slots 1" "--res 1920x1080 --preset medium --frames 96 --slots 48" "--res 3840x2160 --preset veryslow --frames 2 --slots 1" "--res 3840x2160 --preset veryslow --frames 24 --slots 24"; do
  timeout 300 python tools/ctu_devbench.py $args >> $OUT 2>&1
done
cat $OUT | head -250
