#!/bin/bash
# Is the SASS of the member kernels in two builds of libgreb_b200.so the same instruction stream?
#   tools/sass_diff.sh <old.so> [new.so]      (new defaults to the in-tree library)
# Used to state that an ncu capture of an earlier commit still describes HEAD (profiles/README.md):
#   git archive <commit> greb-climate-model_b200/csrc include | tar -x -C /tmp/old
#   make -C /tmp/old/greb-climate-model_b200/csrc OUT=/tmp/old/lib_old.so GRID=/tmp/old/grid_old.so
old=$1
new=${2:-$(dirname "$0")/../greb-climate-model_b200/libgreb_b200.so}
ext() {
  cuobjdump -sass "$1" | awk -v k="$2" '/Function : /{p=(index($3,k)>0)} p' | grep -E "^\s+/\*[0-9a-f]{4,}\*/" |
    sed -E 's#^\s+/\*[0-9a-f]+\*/\s+##; s#\s*/\* 0x[0-9a-f]+ \*/##'
}
rc=0
for k in greb_member_kernelILi0ELi0 greb_member_kernelILi1ELi0 greb_member_kernelILi0ELi1 greb_member_kernelILi1ELi1 \
         greb_circulation_kernelILi0 greb_circulation_kernelILi1; do
  a=$(mktemp) b=$(mktemp)
  ext "$old" $k > $a
  ext "$new" $k > $b
  d=$(diff $a $b | grep -c "^[<>]")
  echo "$k: $(wc -l < $a) / $(wc -l < $b) instructions, $d differing lines"
  [ "$d" != 0 ] && rc=1
  rm -f $a $b
done
exit $rc
