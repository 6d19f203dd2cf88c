for align in 16 128; do for u in 4 8; do
echo "== align=$align U=$u"; DDRL_ROW_ALIGN=$align DDRL_GATHER_U=$u python tools/micro_replay.py 2>&1 | grep -E "'C2'.*sample.*2048|'C3'.*sample.*256|'C1'.*sample.*2048|'C2'.*store.*65536|'C3'.*store.*1048576" | cut -c1-140
done; done
