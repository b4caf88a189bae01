#!/bin/bash
# One gpurun call of the build -> measure loop: new-path tests, bench lines, ncu launch list and full capture.
#   usage: tools/gpu_round.sh <tag> [what ...]     what in: tests fulltests c2 c5 ncu
tag=$1; shift
what="$*"; [ -z "$what" ] && what="tests c2 c5 ncu"
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests) timeout 900 python -m pytest tests/test_fused_iteration.py -m gpu -x -q > gpurun_out/${tag}_fused_tests.log 2>&1; tail -3 gpurun_out/${tag}_fused_tests.log;;
    fulltests) timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; tail -3 gpurun_out/${tag}_gputests.log;;
    c2) timeout 600 python bench.py --workload C2 --steps 10 --no-grad --no-cpu > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; cat gpurun_out/${tag}_bench_c2.json | head -c 3000; tail -3 gpurun_out/${tag}_bench_c2.err;;
    c5) timeout 900 python bench.py --steps 10 --no-cpu > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err; cat gpurun_out/${tag}_bench_c5.json | head -c 4000; tail -3 gpurun_out/${tag}_bench_c5.err;;
    c5nospec) BN_B200_SPEC_FILTER=0 timeout 900 python bench.py --steps 10 --no-cpu --no-grad > gpurun_out/${tag}_bench_c5_nospec.json 2> gpurun_out/${tag}_bench_c5_nospec.err; cat gpurun_out/${tag}_bench_c5_nospec.json | head -c 1500; tail -3 gpurun_out/${tag}_bench_c5_nospec.err;;
    c5u4) BN_B200_LIBNAME=libbn_u4.so timeout 900 python bench.py --steps 10 --no-cpu --no-grad > gpurun_out/${tag}_bench_c5_u4.json 2> gpurun_out/${tag}_bench_c5_u4.err; cat gpurun_out/${tag}_bench_c5_u4.json | head -c 1500; tail -3 gpurun_out/${tag}_bench_c5_u4.err;;
    c5nolin) BN_B200_LINEAR_POST=0 timeout 900 python bench.py --steps 10 --no-cpu --no-grad > gpurun_out/${tag}_bench_c5_nolin.json 2> gpurun_out/${tag}_bench_c5_nolin.err; cat gpurun_out/${tag}_bench_c5_nolin.json | head -c 1500; tail -3 gpurun_out/${tag}_bench_c5_nolin.err;;
    reftests) timeout 900 python -m pytest tests/test_reference_golden.py -m gpu -q > gpurun_out/${tag}_ref_tests.log 2>&1; tail -15 gpurun_out/${tag}_ref_tests.log;;
    mgputests) timeout 1500 python -m pytest tests/test_distributed_gpu.py tests/test_latent_sharding.py -m gpu -q > gpurun_out/${tag}_mgpu_tests_n${NG}.log 2>&1; tail -8 gpurun_out/${tag}_mgpu_tests_n${NG}.log;;
    mgpubench) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG} --steps 10 --no-cpu --no-grad > gpurun_out/${tag}_bench_c5_n${NG}.json 2> gpurun_out/${tag}_bench_c5_n${NG}.err; tail -c 3000 gpurun_out/${tag}_bench_c5_n${NG}.json; tail -5 gpurun_out/${tag}_bench_c5_n${NG}.err;;
    c1) timeout 600 python bench.py --workload C1 --steps 50 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err; head -c 2500 gpurun_out/${tag}_bench_c1.json; tail -3 gpurun_out/${tag}_bench_c1.err;;
    c3) timeout 900 python bench.py --workload C3 --steps 10 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; head -c 2500 gpurun_out/${tag}_bench_c3.json; tail -3 gpurun_out/${tag}_bench_c3.err;;
    c4) timeout 1500 python bench.py --workload C4 --steps 2 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; head -c 3500 gpurun_out/${tag}_bench_c4.json; tail -3 gpurun_out/${tag}_bench_c4.err;;
    c5full) timeout 1200 python bench.py > gpurun_out/${tag}_bench_c5_default.json 2> gpurun_out/${tag}_bench_c5_default.err; head -c 5000 gpurun_out/${tag}_bench_c5_default.json; tail -3 gpurun_out/${tag}_bench_c5_default.err;;
    refarm) timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; head -c 2500 gpurun_out/${tag}_bench_ref.json; tail -3 gpurun_out/${tag}_bench_ref.err;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python tools/prof_iter.py 10000000 3 > gpurun_out/${tag}_under_ncu.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'it_(reduce|filter|smooth)' -s 5 -c 5 -o gpurun_out/${tag}_c2_full python tools/prof_iter.py 10000000 2 > gpurun_out/${tag}_ncu.log 2>&1
      ncu -i gpurun_out/${tag}_c2_full.ncu-rep --page raw --csv > gpurun_out/${tag}_c2_raw.csv 2>/dev/null
      tail -2 gpurun_out/${tag}_ncu.log;;
    ncu5)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_c5_launches.csv python tools/prof_iter.py 100000000 2 > gpurun_out/${tag}_c5_under_ncu.log 2>&1
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'it_(reduce|filter|smooth)' -s 6 -c 6 -o gpurun_out/${tag}_c5_full python tools/prof_iter.py 100000000 2 > gpurun_out/${tag}_c5_ncu.log 2>&1
      ncu -i gpurun_out/${tag}_c5_full.ncu-rep --page raw --csv > gpurun_out/${tag}_c5_raw.csv 2>/dev/null
      tail -2 gpurun_out/${tag}_c5_ncu.log;;
    list12)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_n12m_launches.csv python tools/prof_iter.py 12500000 3 > gpurun_out/${tag}_n12m_under_ncu.log 2>&1; tail -2 gpurun_out/${tag}_n12m_under_ncu.log;;
    fp32) timeout 900 python -m pytest tests/test_fp32_mode.py -m gpu -q -x > gpurun_out/${tag}_fp32_tests.log 2>&1; tail -15 gpurun_out/${tag}_fp32_tests.log
      timeout 600 python tools/bench_fp32.py 10000000 10 > gpurun_out/${tag}_fp32_n1e7.json 2> gpurun_out/${tag}_fp32_n1e7.err; cat gpurun_out/${tag}_fp32_n1e7.json; tail -3 gpurun_out/${tag}_fp32_n1e7.err
      timeout 600 python tools/bench_fp32.py 100000000 5 > gpurun_out/${tag}_fp32_n1e8.json 2> gpurun_out/${tag}_fp32_n1e8.err; cat gpurun_out/${tag}_fp32_n1e8.json; tail -3 gpurun_out/${tag}_fp32_n1e8.err;;
    racecheck)
      for tool in racecheck synccheck memcheck; do
        timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/racecheck_small.py > gpurun_out/${tag}_sanitizer_${tool}.log 2>&1
        tail -4 gpurun_out/${tag}_sanitizer_${tool}.log
      done;;
    gd) timeout 900 python -m pytest tests/test_small_d_generic.py tests/test_spacetime.py -m gpu -q > gpurun_out/${tag}_gd_st_tests.log 2>&1; tail -5 gpurun_out/${tag}_gd_st_tests.log;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log;;
    smalld) timeout 600 python tools/bench_small_d.py 200000 > gpurun_out/${tag}_small_d.json 2> gpurun_out/${tag}_small_d.err; cat gpurun_out/${tag}_small_d.json; tail -3 gpurun_out/${tag}_small_d.err;;
  esac
done
