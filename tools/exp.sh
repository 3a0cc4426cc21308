run() { # lib tb cb nbuf
  QVNT_B200_LIB=$1 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --tile-bits $2 --chunk-bits $3 --nbuf $4 > gpurun_out/b.json 2> gpurun_out/b.err
  python -c "
import json;d=json.load(open('gpurun_out/b.json'));print('$1'[-14:], $2, $3, $4, round(d['value']), round(d['ms_per_step']), round(d['roofline']['avg_launch_ms'],3), round(d['roofline']['frac'],3), d['roofline']['passes_per_step'])"
}
A=$PWD/qvnt_b200/libqvnt_b200.so
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run $A 12 7 0
run $A 12 4 0
run $A 11 4 1
run $A 11 4 2
python tools/tile_probe.py --no-mem --tile-bits 12 | grep -E "^---|^rx|^cx|^h_|^rz"
