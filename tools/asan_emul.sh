#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over every kernel body and the whole host side of the C ABI:
# the host-emulation build (g++ -DVMX_HOST_EMUL: the kernels' bodies run thread by thread on the CPU, "device"
# memory is the heap) instrumented and driven by the CPU parity tests.  compute-sanitizer is closed on the GPU
# pool of this project (profiles/r04_sanitizers.txt), so this is the memory checker of the index arithmetic the
# CUDA build shares with the emulation build; the CUDA-only paths (warp-cooperative kernels, shared-memory staging)
# are covered by the bit-exact GPU parity suite instead.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/asan
( cd verificatum-vmn_b200/csrc && g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined \
    -fno-sanitize-recover=undefined -DVMX_HOST_EMUL -x c++ -fPIC -shared -o ../../build/asan/libvmx_emul_asan.so vmx.cu )
( cd verificatum-vmn_b200/csrc && g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined \
    -fno-sanitize-recover=undefined -fPIC -shared -o ../../build/asan/libvmnv_asan.so vmnv_native.cpp -ldl -lcrypto -lpthread )
export VMX_EMUL_LIBRARY=$PWD/build/asan/libvmx_emul_asan.so
export VMNV_LIBRARY_PATH=$PWD/build/asan/libvmnv_asan.so
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
python -m pytest tests/test_engine_emul.py tests/test_vmnv_native.py tests/test_abi_misuse.py -x -q -p no:cacheprovider "$@"
