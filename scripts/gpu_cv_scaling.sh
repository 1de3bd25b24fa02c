N=${1:-8}
cd summarizer_b200
S=splits/tvsum_splits.json,splits/summe_splits.json
t0=$(date +%s.%N)
if [ "$N" = "1" ]; then
  timeout 900 python main.py -m vasnet -s $S -c yes -e 50 -t 10 > ../gpurun_out/cv_n$N.log 2>&1
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29851 main.py -m vasnet -s $S -c yes -e 50 -t 10 > ../gpurun_out/cv_n$N.log 2>&1
fi
t1=$(date +%s.%N)
echo "{\"config\": \"VASNet 5-fold cross-validation on both synthetic datasets (10 fold jobs), 50 epochs, fold-parallel\", \"n_gpus\": $N, \"wall_s\": $(python -c "print(round($t1-$t0,2))")}" | tee ../gpurun_out/cv_n$N.json
grep -E "Cross-validation" ../gpurun_out/cv_n$N.log | sort -u | tail -4
