# Round 2, GPU call AE: scatter dgrad over several N tiles (resnet8_u64's last layer); training tests, u32 / u64 step
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
timeout 200 python bench.py --steps 3 --extras cfg4,cfg4u64 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
