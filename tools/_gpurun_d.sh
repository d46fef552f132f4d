MARL_B200_SO=$PWD/codebase_b200/csrc/libmarlb200_ts.so python tools/tc_fwd_micro.py 26624 1 2>&1 | grep "^TS" | tail -3
