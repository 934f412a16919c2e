# one live stream through the command line: wall time per block for BASELINE configs[0] (mode 0 mono, ~10 s) and stereo
python - <<'PY'
import sys; sys.path.insert(0, ".")
import dy4_b200
m = dy4_b200.mode_params(0)
dy4_b200.synth.make_stream(0, 468 * m.block_size // 2, 65).tofile("/tmp/stream.raw")
PY
for ch in mono stereo; do for per in 1 8; do
  s=$(date +%s.%N); ./3dy4-real-time-software-defined-radio-_b200/dy4_project 0 $ch $per < /tmp/stream.raw > /tmp/out.pcm 2>/dev/null; e=$(date +%s.%N)
  python -c "import os; n=468; t=$e-$s; print('dy4_project 0 $ch blocks_per_call=$per: %.3f s for %d blocks (9.98 s of signal): %.2f ms per 21.33 ms block incl. process start-up, output %d bytes' % (t, n, 1e3*t/n, os.path.getsize('/tmp/out.pcm')))"
done; done
