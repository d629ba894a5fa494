# prints the name:lo-hi phase ranges of gi_lane.cuh for tools/ncu_by_phase.py / sass_size_by_phase.py
L=$(dirname "$0")/../motion-generation-using-quadratic-programs_b200/csrc/gi_lane.cuh
ln() { grep -n "$1" $L | head -1 | cut -d: -f1; }
g0=$(ln "MGQP_DEV double ln_distance"); wk=$(ln "^struct LaneWalk"); sA=$(ln "^MGQP_DEV void ln_stage_G"); cb=$(ln "^MGQP_DEV unsigned ln_column_byte"); t0=$(ln "^MGQP_DEV void gi_lane_tile")
pA=$(ln "phase A: stage G, g0"); su=$(ln "set-up$"); pB=$(ln "phase B: inequality columns"); jl=$(ln "J moves into REGISTERS"); lp=$(ln "One pass of this loop"); sc=$(ln "step 1 / 2: scan"); d2=$(ln "step 2a"); s2=$(ln "step 2b"); ad=$(ln "add_constraint (App"); de=$(ln "delete_constraint(del_l)"); dg=$(ln "if (degenerate) {")
echo givens:$g0-$((wk-4)) walk:$((wk-3))-$((sA-3)) stageG:$((sA-2))-$((cb-6)) stageCI:$((cb-5))-$((t0-3)) hdr:$((t0-2))-$((pA-1)) phaseA:$pA-$((su-1)) setup:$su-$((pB-1)) phaseB:$pB-$((jl-1)) init:$jl-$((sc-8)) scan:$((sc-7))-$((d2-9)) direction:$((d2-8))-$((s2-1)) step:$s2-$((ad-1)) add:$ad-$((de-1)) delete:$de-$((dg-1)) tail:$dg-$((dg+40))
