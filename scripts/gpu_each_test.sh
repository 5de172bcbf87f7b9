for t in test_search_parity_random test_search_empty_and_single test_sga_many_seeds_slow_path test_sw_parity_acgt test_sw_parity_iupac_and_n test_sw_parity_lengths_and_strip_edges test_sw_literal_small test_sw_empty test_align_kats test_align_parity_c1 test_align_parity_n_reads_and_edges test_align_parity_150_250 test_align_batch_split_invariance test_synthetic_index_parity_midsize; do
  echo "=== $t"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$t" 2>&1 | grep -E "^E |passed|failed|Error" | head -12
done
