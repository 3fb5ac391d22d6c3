for k in auto 0 128 256 320 384 408; do
  if [ $k = auto ]; then a=""; else a="--hot-nodes $k"; fi
  python bench.py --net none --no-cpu-baseline --no-e2e $a > gpurun_out/sw_$k.json 2>gpurun_out/sw_$k.err || tail -3 gpurun_out/sw_$k.err
  python -c "
import json; d=json.load(open('gpurun_out/sw_$k.json')); print('K=$k', 'sims/s %.3e'%d['value'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
done
python bench.py --net none --no-cpu-baseline --no-e2e --lanes 32 > gpurun_out/sw_l32.json; python -c "
import json; d=json.load(open('gpurun_out/sw_l32.json')); print('lanes32', 'sims/s %.3e'%d['value'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'])"
python bench.py --net none --no-cpu-baseline --no-e2e --games 16384 > gpurun_out/sw_16k.json; python -c "
import json; d=json.load(open('gpurun_out/sw_16k.json')); print('E=16384', 'sims/s %.3e'%d['value'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
python bench.py --net none --no-cpu-baseline --no-e2e --games 16384 --sims 800 --steps 10 > gpurun_out/sw_16k800.json; python -c "
import json; d=json.load(open('gpurun_out/sw_16k800.json')); print('E=16384 S=800', 'sims/s %.3e'%d['value'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
