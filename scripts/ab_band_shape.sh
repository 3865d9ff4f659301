cd /root/repo
python scripts/agg_time_escalated.py 608x7680x512 3
for v in w0 w4 old055; do
  SMT_B200_LIB=$PWD/stereo_match_traditional_b200/csrc/build_var/$v/libsmt_b200.so python scripts/agg_time_escalated.py 608x7680x512 3
done
