# compute-sanitizer racecheck of the warp-specialised rollout, product build vs the -DBRL_PLAIN_EPISODE_LOAD debug build
# (brl_b200/lib/libbrl_plainload.so, built beforehand: python scripts/build_variant.py plainload -DBRL_PLAIN_EPISODE_LOAD)
O=gpurun_out/r2h; mkdir -p $O
python scripts/rollout_checksum.py > $O/checksum_product.txt 2>&1
BRL_B200_LIB=brl_b200/lib/libbrl_plainload.so python scripts/rollout_checksum.py > $O/checksum_plainload.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python scripts/rollout_checksum.py > $O/racecheck_product.txt 2>&1
BRL_B200_LIB=brl_b200/lib/libbrl_plainload.so timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python scripts/rollout_checksum.py > $O/racecheck_plainload.txt 2>&1
for f in checksum_product checksum_plainload; do echo $f; cat $O/$f.txt | tail -1; done
for f in racecheck_product racecheck_plainload; do echo $f; grep -c "Race reported" $O/$f.txt; grep "RACECHECK SUMMARY\|checksum" $O/$f.txt; done
