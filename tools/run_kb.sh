for b in lock6_s2 lock6_s3 spec6_s2 spec6_s3 spec6_bar_s2 f3_s2 f3_s3 f1_s2; do echo "== $b"; ./tools/kbench_$b 4000000 1000 10 | head -1; ./tools/kbench_$b 400000 10000 10 | head -1; done
