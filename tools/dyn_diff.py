import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / 'oracle', ROOT / 'tests'): sys.path.insert(0, str(p))
import numpy as np, torch
from helpers import load_golden, schema_for, actions_of
from citylearn_b200 import CityLearnEnv, schema as S
from citylearn_oracle import OracleEnv
z, cfg, _ = load_golden('c3_marl')
sch, src, ov = schema_for(cfg)
env = CityLearnEnv(sch, data_source=src, num_envs=1, **ov)
orc = OracleEnv(env.spec, 1); orc.reset(); env.reset()
names = list(S.DYN)
acts = actions_of(z)[0]
worst = {}
for k in range(len(acts)):
    env.step(acts[k][None]); _, _, od, odyn = orc.step(acts[k][None])
    tr = env.trace[0].cpu().numpy().astype('float64')
    d = np.abs(tr - odyn[0]); 
    for j, n in enumerate(names):
        m = np.nanmax(d[:, j])
        if m > worst.get(n, (0, 0))[0]: worst[n] = (float(m), k)
    dd = np.abs(env.district[0].cpu().numpy() - od[0]).max()
    if dd > worst.get('district', (0, 0))[0]: worst['district'] = (float(dd), k)
for n, v in worst.items(): print(n, v)
h_d = env._hist_dyn[:env.time_step].cpu().numpy()
print('hist vs last trace equal', np.array_equal(h_d[-1], env.trace[0].cpu().numpy(), equal_nan=True), env.time_step)
from citylearn_b200.evaluate import History, evaluate
df = env.evaluate()
got = {(r['name'], r['cost_function']): r['value'] for r in df.to_dict('records')}
ref = {(r['name'], r['cost_function']): r['value'] for r in json.loads(bytes(z['evaluate']).decode())}
for k in ref:
    g, v = got[k], ref[k]
    if v is not None and g is not None and abs(g - v) > 2e-6 * max(1, abs(v)): print('KPI', k, g, v)
print('outage shapes', env._outage.shape, orc.outage.shape, np.array_equal(env._outage, orc.outage))
hd = env._hist_district[:env.time_step].cpu().numpy()
print('district hist len', hd.shape, 'max', hd[:, 0].max(), 'argmax', hd[:, 0].argmax())
