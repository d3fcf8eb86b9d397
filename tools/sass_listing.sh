#!/bin/bash
# SASS of the kernels the bench runs (TMA instances of the five extractor kernels and the tcgen05 all-pairs kernel), without the encoding
# columns, plus a count of the mnemonics that prove TMA / tcgen05 (B200_PROFILING.md) -> profiles/rNN_sass_*.txt
set -e
cd "$(dirname "$0")/.."
R=${1:-r02}
LIB=monoorbslam3_b200/lib/liborbfe.so
strip_enc() { grep -vE "^(Fatbin|=====|arch =|code version|host =|compile_size|identifier|[[:space:]]*code for)" | sed -E 's#/\*[^*]*\*/##g; s#[[:space:]]+$##; s#^[[:space:]]{8,}#    #' | awk 'NF'; }
{
for f in _ZN5orbfe8k_resizeILb1ELb1EEEv14CUtensorMap_stNS_10ResizeArgsE _ZN5orbfe13k_fast_planesILb1EEEvNS_8LevelSetENS_7TmapSetENS_9Fast2ArgsE \
         _ZN5orbfe8k_octreeILi512EEEvNS_8LevelSetENS_7OctArgsE _ZN5orbfe6k_blurILb1EEEvNS_8LevelSetENS_7TmapSetENS_8BlurArgsE \
         _ZN5orbfe10k_describeILb1EEEvNS_8LevelSetENS_9PatchMapsENS_8DescArgsE; do
    cuobjdump -sass -fun "$f" $LIB 2>/dev/null | strip_enc
done
} > profiles/${R}_sass_extractor.txt
cuobjdump -sass -fun _ZN5orbfe13k_allpairs_tcE14CUtensorMap_stS0_NS_6TcArgsE $LIB 2>/dev/null | strip_enc > profiles/${R}_sass_allpairs_tc.txt
{
echo "mnemonic counts (cuobjdump -sass of $LIB)"
for k in extractor allpairs_tc; do
    echo "== profiles/${R}_sass_$k.txt"
    for m in UTMALDG UTMASTG UTCIMMA UTCHMMA UTCBAR LDTM STTM SYNCS VABSDIFF4 VIMNMX3 IDP IMMA; do
        printf "%-10s %d\n" $m $(grep -c "[[:space:]]$m" profiles/${R}_sass_$k.txt || true)
    done
done
} > profiles/${R}_sass_counts.txt
cat profiles/${R}_sass_counts.txt
