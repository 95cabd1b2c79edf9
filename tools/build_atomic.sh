#!/bin/bash
# builds the product libraries into /tmp and moves them into place atomically (a gpurun snapshot may be taken at any time)
set -e
cd /root/repo/openh264_b200/csrc
rm -f /tmp/libopenh264_b200.so.new
make -s -j8 OUT=/tmp/libopenh264_b200.so.new > /tmp/build_atomic.log 2>&1 || { grep -B2 -A8 "error" /tmp/build_atomic.log | head -60; echo BUILD FAILED; exit 1; }
test -f /tmp/libopenh264_b200.so.new
cp /tmp/libopenh264_b200.so.new ../libopenh264_b200.so.tmp && mv ../libopenh264_b200.so.tmp ../libopenh264_b200.so
if [ -d /root/reference/codec ]; then
  cd ../wels
  SRCS=$(ls *.cpp)
  g++ -O2 -std=c++17 -fPIC -shared -Wall -I/root/reference/codec/api/wels -I../../include -I/usr/local/cuda/include $SRCS -o /tmp/libopenh264_b200_wels.so.new \
      -L.. -lopenh264_b200 -L/usr/local/cuda/lib64 -lcudart -lpthread -Wl,-rpath,'$ORIGIN' -Wl,-soname,libopenh264_b200_wels.so
  cp /tmp/libopenh264_b200_wels.so.new ../libopenh264_b200_wels.so.tmp && mv ../libopenh264_b200_wels.so.tmp ../libopenh264_b200_wels.so
fi
echo built
