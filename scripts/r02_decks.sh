# round 2: the shipped dust benchmark decks to convergence on the device (fixtures of tests/golden), files kept
for name in p0tau1 p0tau10 p0tau100 tau1.000; do
  out=gpurun_out/r02_deck_$name
  mkdir -p $out
  timeout 600 python scripts/run_dust_deck.py --golden tests/golden/deck_$name.npz --out $out > gpurun_out/r02_dust_deck_${name}_b200.jsonl 2> $out/stderr.log

  tail -2 gpurun_out/r02_dust_deck_${name}_b200.jsonl | cut -c1-400
  rm -f $out/grid0.out $out/dustGrid.out          # MBs of per-cell text; SED/summary/tauNu/photoSource are kept
done
