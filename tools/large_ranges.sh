# name:lo-hi phase ranges of gi_large.cuh for tools/ncu_by_phase.py
L=$(dirname "$0")/../motion-generation-using-quadratic-programs_b200/csrc/gi_large.cuh
ln() { grep -n "$1" $L | head -1 | cut -d: -f1; }
am=$(ln "auto block_argmin"); us=$(ln "auto upper_solve"); su=$(ln "set-up$"); di=$(ln "auto direction"); ac=$(ln "auto add_constraint"); dc=$(ln "auto delete_constraint"); eq=$(ln "equality constraints$"); el=$(ln "ELL form of CI: one pass"); lp=$(ln "while (!done) {"); sl=$(ln "// step lengths"); en=$(wc -l < $L)
echo reduce:$am-$((us-1)) upper_solve:$us-$((su-1)) setup:$su-$((di-1)) direction:$di-$((ac-1)) add:$ac-$((dc-1)) delete:$dc-$((eq-1)) equality:$eq-$((el-3)) ellbuild:$((el-2))-$((lp-1)) scan:$lp-$((sl-7)) step:$((sl-6))-$en
