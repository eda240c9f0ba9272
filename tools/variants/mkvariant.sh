#!/bin/bash
# usage: mkvariant.sh <name> <python-edit-script>   (edits a temp copy of csrc, builds cotr_b200/lib/ab/<name>.so)
name=$1; script=$2
T=$(mktemp -d)
mkdir -p $T/cotr_b200
cp -r /root/repo/cotr_b200/csrc $T/cotr_b200/csrc
cp -r /root/repo/include $T/include
python3 $script $T/cotr_b200/csrc || exit 1
cd $T/cotr_b200/csrc
pids=""
for f in *.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -c $f -o ${f%.cu}.o 2>$f.log & pids="$pids $!"; done
fail=0
for p in $pids; do wait $p || fail=1; done
if [ $fail = 1 ]; then grep -h "error" *.log | head; cd /root/repo; rm -rf $T; exit 1; fi
mkdir -p /root/repo/cotr_b200/lib/ab
nvcc -shared -cudart static -o /root/repo/cotr_b200/lib/ab/$name.so *.o 2>/dev/null
cd /root/repo; rm -rf $T
ls -la /root/repo/cotr_b200/lib/ab/$name.so
