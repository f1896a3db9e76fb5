#!/bin/bash
# SASS listings of the hot kernels from the library as built at HEAD (cuobjdump works without a GPU): bash profiles/sass.sh r02
R=${1:-r02}
LIB=qhg4_b200/libqhg_b200.so
names=$(cuobjdump -elf $LIB 2>/dev/null | grep -o "_ZN3qhg[A-Za-z0-9_]*" | grep -v "_param_" | sort -u)
pick() { echo "$names" | grep "$1" | head -1; }
dump() {  # dump <symbol> <file> <what>
  { echo "# $3"; echo "# symbol $1, library built from $(git rev-parse --short HEAD) (+ working tree), nvcc $(nvcc --version | grep -o 'release [0-9.]*')";
    cuobjdump -sass -fun "$1" $LIB | grep -v "^\s*/\* 0x" ; } > "$2"
  echo "$2: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' $2) instructions"
}
dump "$(pick 'k_seg_decideILb1ELi4ELb0ELb0E')" profiles/sass_${R}_k_seg_decide_tut5_sb4.txt "k_seg_decide<SPEC=true, SB=4>: pass 1 of the fast path, tutorial action order, dense populations (the bench's C4 kernel)"
dump "$(pick 'k_seg_decideILb1ELi16ELb0ELb0E')" profiles/sass_${R}_k_seg_decide_tut5_sb16.txt "k_seg_decide<SPEC=true, SB=16>: the same for sparse populations (C2: below 32 agents per cell; SB=8 below 64)"
dump "$(pick 'k_seg_decideILb0ELi8ELb1ELb1E')" profiles/sass_${R}_k_seg_decide_gen_nav_sb8.txt "k_seg_decide<SPEC=false, SB=8, GEN, NAV>: interpreted program, Genetics + Navigate (C5)"
dump "$(pick 'k_cell_scatterILb0ELi384ELi6ELi4ELi1ELb0E')" profiles/sass_${R}_k_cell_scatter_dense.txt "k_cell_scatter<GEN=false, SCH=384, 6 CTAs per SM, 4 cells per grab, 1 stage>: pass 2, dense populations (UBLKCP = cp.async.bulk, SYNCS = mbarrier, MATCH = the movers of one (cell, direction))"
dump "$(pick 'k_cell_scatterILb1ELi256ELi8ELi16ELi1ELb0E')" profiles/sass_${R}_k_cell_scatter_genetic_sparse.txt "k_cell_scatter<GEN=true, SCH=256, 8 CTAs per SM, 16 cells per grab>: pass 2 with genome handles, birth records and genome rows for the migrants"
dump "$(pick 'k_make_offspring')" profiles/sass_${R}_k_make_offspring.txt "k_make_offspring: Genetics::makeOffspring, one warp per birth"
