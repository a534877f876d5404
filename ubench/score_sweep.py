"""Brute-force scoring launch efficiency vs cloud size and hypothesis count (run through gpurun).

    python ubench/score_sweep.py            # runs itself once per PR_SCORE_GEOM setting
"""
import os, subprocess, sys
sys.path.insert(0, '.')

def child():
    import dialog_b200 as D
    from dialog_b200 import synth
    pr = D.PlaneRansac(0)
    peak = pr.measure_ffma_peak()
    sc = synth.indoor_scene()
    for n in (1_000_000, 2_000_000, 3_000_000, 5_000_000, 10_000_000):
        pts = sc.points(0, n)
        pr.set_cloud(pts)
        for K in (1024, 3072, 4096):
            tri = D.host_draw_triples(n, K)
            pr.score(tri, 0.1); pr.score(tri, 0.1)
            pr.profile_enable(True); pr.profile_reset()
            for _ in range(5): c = pr.score(tri, 0.1)
            p = pr.profile(); pr.profile_enable(False)
            tf = 6.0 * p.pairs_scored / (p.ms_score * 1e-3) / 1e12
            print(f"GEOM={os.environ.get('PR_SCORE_GEOM','1')} N={n:>9} K={K:>5}: {p.ms_score/5:7.3f} ms  {p.launches_score//5} launch(es)  "
                  f"{100*tf/peak:5.1f}% of {peak:.1f} TF  checksum {int(c.sum())}", flush=True)

if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for g in ('0', '1'):
            subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, PR_SCORE_GEOM=g), check=True)
