mkdir -p gpurun_out/r2
SEL="test_topic_widths or test_processes_concurrent or test_transform_and_perplexity_parity or test_transform_pipeline_ranges or test_k256_large or test_burn_in or test_transform_pass_counts or test_empty_documents or test_step_issued"
for g in 2 4; do
  LDA_B200_DOCS_PER_WARP_FIT=$g LDA_B200_DOCS_PER_WARP_TRANSFORM=$g python -m pytest tests/test_lda_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/r2/gk_tests_$g.log 2>&1; echo "gd=$g tests: $(tail -1 gpurun_out/r2/gk_tests_$g.log)"
done
for g in 1 2 4; do
  LDA_B200_DOCS_PER_WARP_TRANSFORM=$g python bench.py --workload C5 --docs 524288 --steps 2 --warmup 1 > gpurun_out/r2/gk_c5_$g.json 2>/dev/null
  LDA_B200_DOCS_PER_WARP_FIT=$g python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-parity --skip-extras > gpurun_out/r2/gk_c3_$g.json 2>/dev/null
done
python - <<PY
import json
for g in (1,2,4):
    a=json.load(open("gpurun_out/r2/gk_c5_%d.json"%g)); b=json.load(open("gpurun_out/r2/gk_c3_%d.json"%g))
    print("docs/warp %d: C5 524k docs %.1f ms/call %.3e upd/s | C3 fit %.2f ms/iter %.3e, kernel/floor %.3f, clocks %s" % (g, a["ms_per_step"], a["value"], b["ms_per_step"], b["value"], b["roofline"]["memory_floor"]["kernel_frac_of_floor"], b["clocks"]["sm_mhz"]))
PY
