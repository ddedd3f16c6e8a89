#!/bin/bash
# GPU box: identity of the CTU driver (quick set) + device-only throughput and phase profile.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /dev/shm/c
OUT=gpurun_out/step.log
: > $OUT
REF=oracle/_ref/kvazaar; CTU=oracle/_ref/kvazaar_ctu; LIB=$PWD/kvazaar_b200/libkvzcuda.so
gen() { [ -f /dev/shm/c/$1.yuv ] || python tools/synth_yuv.py $2 $3 $4 /dev/shm/c/$1.yuv $5; }
fps() { grep -E "FPS" $1 | tr '\n' ' '; }
idrun() { # name w h preset qp frames
  $REF -i /dev/shm/c/$1.yuv --input-res $2x$3 -o /dev/shm/c/ref.hevc --preset $4 -q $5 -p 1 -n $6 2>/dev/shm/c/ref.err
  KVZ_CTU_PROVIDER=$LIB timeout 600 $CTU -i /dev/shm/c/$1.yuv --input-res $2x$3 -o /dev/shm/c/ctu.hevc --preset $4 -q $5 -p 1 -n $6 ${@:7} 2>/dev/shm/c/ctu.err
  local same=DIFFERENT; cmp -s /dev/shm/c/ref.hevc /dev/shm/c/ctu.hevc && same=IDENTICAL
  echo "$same $* ref: $(fps /dev/shm/c/ref.err) ctu: $(fps /dev/shm/c/ctu.err) $(grep -c MISMATCH /dev/shm/c/ctu.err)" >> $OUT
}
gen a264 264 200 2
gen n264 264 200 2 --noisy
gen a1080 1920 1080 48
gen a2160 3840 2160 16
idrun a264 264 200 veryslow 22 2
idrun n264 264 200 veryslow 22 2
idrun n264 264 200 medium 27 2
idrun a1080 1920 1080 medium 27 48
idrun a1080 1920 1080 medium 27 48 --owf 40
idrun a2160 3840 2160 veryslow 22 16
idrun a2160 3840 2160 veryslow 22 16 --owf 15
for args in "--res 1920x1080 --preset medium --frames 16 --slots 1" "--res 1920x1080 --preset medium --frames 48 --slots 16" "--res 1920x1080 --preset medium --frames 64 --slots 48" \
            "--res 3840x2160 --preset veryslow --frames 4 --slots 1" "--res 3840x2160 --preset veryslow --frames 16 --slots 16"; do
  timeout 900 python tools/ctu_devbench.py $args >> $OUT 2>&1
done
cat $OUT
