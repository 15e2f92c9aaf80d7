bash tools/run_named.sh 8
