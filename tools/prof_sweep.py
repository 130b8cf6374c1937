"""cProfile of pruner.prune() on the full-size BLIP-2 (host-side hot spots of the stage-2 sweep)."""
import cProfile, pstats, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import prune_wall
which = sys.argv[1:] or ["sparsegpt"]
pr = cProfile.Profile()
pr.enable()
out = prune_wall(which)
pr.disable()
print(out)
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats("tottime").print_stats(14)
st.print_callers("builtins.compile")
st.sort_stats("cumtime").print_stats(30)
print(s.getvalue()[:14000])
