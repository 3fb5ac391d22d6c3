import json, sys
import numpy as np, torch
sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
run = json.load(open('tests/golden/selfplay_goldens.json'))['data'][2]
print('ref lens', [len(e['samples']) for e in run['episodes']])
gen = az.EpisodeGenerator(model=az.UniformEvaluator(), num_simulations=run['S'], num_episodes=run['E'], game_initial_state=az.Config().sample_initial_state())
np.random.seed(run['seed'])
for b in gen.generate_batches(quota=run['E']):
    print('batch', b.ep_step.tolist(), b.ep_slot.tolist(), b.ep_len.tolist(), b.ep_offset.tolist(), len(b.s_bb0))
np.random.seed(run['seed'])
eps = list(gen.generate_episodes())
print('gen_episodes lens', [len(e) for e in eps])
