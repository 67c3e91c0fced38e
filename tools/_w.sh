MGODPL_WIDE_BVH=1 python tools/gpu_counters.py 2 2>&1 | tail -2
python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['scene']['build_ms']); e=d['extra']; print(e['states_10M_prm_box_h5_v10']['ms'], e['config1_states_10M_sampling_difficulty_box']['ms'], e['config3_goal_sampling_500x1000']['ms'], e['config4_prm_edges_100k_nodes_k10']['ms_edge_validation'], e['config5_orchard_8_trees']['states_10M']['ms'], e['config5_orchard_8_trees']['motions_1M']['ms'])"
