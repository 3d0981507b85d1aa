#!/bin/bash
# e2e of `ftkb200 -f cp --input` on a raw file series: 12 files of the C2 generator, 8192 x 8192 float64 (537 MB each) and
# float32 (268 MB each), read through the CLI's double-buffered page-locked reader.  Usage: scripts/cli_input_timing.sh OUTDIR
OUT=${1:-gpurun_out}
D=/tmp/ftkb_series
mkdir -p $D $OUT
python - <<'PY'
import numpy as np
W = H = 8192
x = np.arange(W, dtype=np.float64)
for k in range(12):
    f = ((x - (4096.3 + 0.1 * k)) ** 2)[None, :] + ((x - (4095.7 + 0.1 * k)) ** 2)[:, None]
    f.tofile(f"/tmp/ftkb_series/me_{k:03d}.f64")
    f.astype(np.float32).tofile(f"/tmp/ftkb_series/me_{k:03d}.f32")
PY
for fmt in float64 float32; do
  ext=f64; [ $fmt == float32 ] && ext=f32
  for rep in 1 2; do    # the second pass reads from the page cache
    ./ftk_b200/bin/ftkb200 -f cp --input "$D/me_%03d.$ext" --input-format $fmt --width 8192 --height 8192 --timesteps 12 --timing \
        --output-type discrete -o $D/out_$ext.txt 2>&1 | tail -3 | sed "s/^/$fmt pass $rep: /" | tee -a $OUT/${TAG:-r02i}_cli_input_timing.log
  done
done
wc -l $D/out_f64.txt $D/out_f32.txt | tee -a $OUT/${TAG:-r02i}_cli_input_timing.log
rm -rf $D
