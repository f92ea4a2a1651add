# developer build of a variant library: bash scripts/dev_build.sh <name> <extra nvcc flags...>
set -e
cd "$(dirname "$0")/../csdotrajectoryplanning_b200/csrc"; mkdir -p _dev/_obj
name=$1; shift
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="$* -O3 -std=c++17 -lineinfo $ARCH -I../../include -I. -Xcompiler -fPIC -ccbin /usr/bin/g++"
$NVCC $COMMON -DCSDO_TU=1 -c dsqp_kernel.cu -o _dev/_obj/dsqp_kernel_short.o &
$NVCC $COMMON -DCSDO_TU=2 -c dsqp_kernel.cu -o _dev/_obj/dsqp_kernel_wide.o &
$NVCC $COMMON -fmad=false -c planes_kernel.cu -o _dev/_obj/planes_kernel.o &
$NVCC $COMMON -c csdo_api.cu -o _dev/_obj/csdo_api.o &
$NVCC $COMMON -c measure_kernel.cu -o _dev/_obj/measure_kernel.o &
wait
$NVCC -shared $ARCH -o _dev/libcsdo_$name.so _dev/_obj/*.o -Xcompiler -fPIC -ccbin /usr/bin/g++ -lcudart
