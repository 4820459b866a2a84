# Build a variant of libmplb.so with extra nvcc flags: tools/build_variant.sh NAME -DMPLB_X=1 ...
# -> mplambda_b200/_variants/libmplb_NAME.so (travels to the GPU box; run with MPLB_LIB=that path)
N=$1; shift
MPLB_NVCC_EXTRA="$*" MPLB_OBJ=/tmp/mplb_obj_$N MPLB_LIB=$(dirname $0)/../mplambda_b200/_variants/libmplb_$N.so python -m mplambda_b200.build --force > /tmp/mplb_build_$N.log 2>&1 && echo built $N || { echo FAILED $N; tail -20 /tmp/mplb_build_$N.log; }
