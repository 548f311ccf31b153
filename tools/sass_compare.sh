#!/bin/bash
# Compare the SASS of every kernel between two build directories, ignoring constant-bank offsets (which move when a
# parameter struct grows). Used to show that a change leaves the shipped kernels untouched when no GPU is at hand:
#   git worktree add /tmp/wt <validated commit> && make -C /tmp/wt/voxel-hashing-sdf_b200 && tools/sass_compare.sh /tmp/wt/voxel-hashing-sdf_b200/build voxel-hashing-sdf_b200/build
# (kernels in anonymous namespaces, e.g. vh_weld.cu, carry a path hash in their names and always show as changed)
A=$1; B=$2
for o in vh_engine vh_alloc vh_integrate vh_mc vh_export vh_map_api vh_shard vh_weld; do
  for side in A B; do
    d=${!side}
    cuobjdump -sass $d/$o.o 2>/dev/null | awk '/Function :/{name=$3} /^ *\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//{print name, $0}' | sed 's#/\* 0x[0-9a-f]* \*/##; s#/\*[0-9a-f]*\*/##; s/\s\+/ /g; s/c\[0x0\]\[0x[0-9a-f]*\]/c[X]/g' > /tmp/sass_$side.txt
  done
  na=$(cut -d' ' -f1 /tmp/sass_A.txt | sort -u | wc -l); nb=$(cut -d' ' -f1 /tmp/sass_B.txt | sort -u | wc -l)
  changed=0
  for k in $(cut -d' ' -f1 /tmp/sass_A.txt | sort -u); do
    if ! diff -q <(grep -F "$k " /tmp/sass_A.txt) <(grep -F "$k " /tmp/sass_B.txt) >/dev/null; then changed=$((changed+1)); echo "  CHANGED: $o $k" | cut -c1-150; fi
  done
  echo "$o: kernels $na -> $nb, changed $changed"
done
