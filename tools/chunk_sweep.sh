for c in 32 64 96 128 192 256; do B200SPEC_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk $c step %.3f'%d['ms_per_step'], ' '.join('%d:%.3f'%(k['frame_size'],k['ms']) for k in d['roofline']['per_kernel']))"; done
