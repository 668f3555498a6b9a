"""Config-size smoke of the waypoint planners (BASELINE configs 2 and 3): a few iterations at full population."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from maaco_path_planing_b200 import blocks_map
from maaco_path_planing_b200.pso import PSOSolver
from maaco_path_planing_b200.ga_solver import GASolver
from maaco_path_planing_b200.mpa import MPA
POL = dict(turn_penalty_factor=0.3, safety_penalty_factor=0.8, min_safe_distance=1.8, diagonal_obstacle_penalty_value=100.0,
           allow_diagonal_moves=True, restrict_diagonal_near_obstacle_policy=True)
g = blocks_map(512, 0.2, seed=3512)
t0 = time.time()
s = PSOSolver(g, num_iterations=2, num_particles=4096, num_waypoints_per_particle=5, w=0.7, c1=1.5, c2=1.5, rng_seed=1, verbose=False, **POL)
res = s.solve(); torch.cuda.synchronize()
print(f'PSO 512^2 N=4096 2 it: {time.time()-t0:.1f}s evals={s.fitness_evaluations} attempts={s.init_attempts} repair_rounds={s.repair_rounds} curve={[round(c,2) for c in s.convergence_curve]} path={len(res[0])}', flush=True)
assert res[0][0] == s.start_node and res[0][-1] == s.target_node
t0 = time.time()
s = GASolver(g, num_generations=2, population_size=4096, num_waypoints_per_chromosome=5, mutation_rate=0.1, crossover_rate=0.8, tournament_size=3, rng_seed=2, verbose=False, **POL)
res = s.solve(); torch.cuda.synchronize()
print(f'GA 512^2 N=4096 2 gen: {time.time()-t0:.1f}s evals={s.fitness_evaluations} attempts={s.init_attempts} curve={[round(c,2) for c in s.convergence_curve]} path={len(res[0])}', flush=True)
assert res[0][0] == s.start_node and res[0][-1] == s.target_node
g2 = blocks_map(100, 0.2, seed=2000)
t0 = time.time()
m = MPA(g2, num_predators=1024, num_iterations=100, FADs_rate=0.2, P_const=0.5, levy_beta=2.0, turn_penalty_factor=0.1, safety_penalty_factor=0.8, min_safe_distance=1.8, diagonal_obstacle_penalty=100.0, rng_seed=3, verbose=False)
res = m.solve_path_planning(); torch.cuda.synchronize()
print(f'MPA 100^2 N=1024 100 it (config 2): {time.time()-t0:.1f}s best={res[5]:.3f} curve0={m.convergence_curve_data[0]:.3f} path={len(res[0])}', flush=True)
assert res[0][0] == m.start_node and res[0][-1] == m.target_node
m2 = MPA(g2, num_predators=1024, num_iterations=100, FADs_rate=0.2, P_const=0.5, levy_beta=2.0, turn_penalty_factor=0.1, safety_penalty_factor=0.8, min_safe_distance=1.8, diagonal_obstacle_penalty=100.0, rng_seed=3, verbose=False)
assert m2.solve_path_planning() == res, 'MPA not deterministic'
print('ok')
