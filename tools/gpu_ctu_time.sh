#!/bin/bash
# GPU box: timing of the reference CLI vs the CLI with the CTU-driver hooks (FPS lines), plus ncu launch lists.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /dev/shm/c
OUT=gpurun_out/ctu_time.log
: > $OUT
REF=oracle/_ref/kvazaar; CTU=oracle/_ref/kvazaar_ctu; LIB=$PWD/kvazaar_b200/libkvzcuda.so
nvidia-smi -L >> $OUT 2>&1
echo "nproc $(nproc)" >> $OUT
gen() { [ -f /dev/shm/c/$1.yuv ] || python tools/synth_yuv.py $2 $3 $4 /dev/shm/c/$1.yuv; }
fps() { grep -E "FPS|Encoding wall" $1 | tr '\n' ' '; }
refrun() { # name w h preset qp frames
  $REF -i /dev/shm/c/$1.yuv --input-res $2x$3 -o /dev/shm/c/ref_$1.hevc --preset $4 -q $5 -p 1 -n $6 2>/dev/shm/c/ref.err
  echo "REF $* : $(fps /dev/shm/c/ref.err)" >> $OUT
}
cturun() { # name w h preset qp frames extra-args... (env passes through)
  local n=$1 w=$2 h=$3 p=$4 q=$5 f=$6; shift 6
  KVZ_CTU_PROVIDER=$LIB timeout 900 $CTU -i /dev/shm/c/$n.yuv --input-res ${w}x$h -o /dev/shm/c/ctu_$n.hevc --preset $p -q $q -p 1 -n $f "$@" 2>/dev/shm/c/ctu.err
  local same=DIFFERENT; cmp -s /dev/shm/c/ref_$n.hevc /dev/shm/c/ctu_$n.hevc && same=IDENTICAL
  echo "CTU $n $p q$q n$f slots=${KVZ_CTU_SLOTS:-8} $* : $same $(fps /dev/shm/c/ctu.err)" >> $OUT
}
gen a1080 1920 1080 48
gen a2160 3840 2160 16
refrun a1080 1920 1080 medium 27 48
cturun a1080 1920 1080 medium 27 48
KVZ_CTU_SLOTS=24 cturun a1080 1920 1080 medium 27 48 --owf 24
refrun a2160 3840 2160 veryslow 22 16
cturun a2160 3840 2160 veryslow 22 16
KVZ_CTU_SLOTS=16 cturun a2160 3840 2160 veryslow 22 16 --owf 16
# launch lists (one frame each)
KVZ_CTU_PROVIDER=$LIB timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ctu_launches_1080p_medium.csv \
  $CTU -i /dev/shm/c/a1080.yuv --input-res 1920x1080 -o /dev/shm/c/x.hevc --preset medium -q 27 -p 1 -n 1 > /dev/null 2>&1
KVZ_CTU_PROVIDER=$LIB timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ctu_launches_2160p_veryslow.csv \
  $CTU -i /dev/shm/c/a2160.yuv --input-res 3840x2160 -o /dev/shm/c/x.hevc --preset veryslow -q 22 -p 1 -n 1 > /dev/null 2>&1
cat $OUT
