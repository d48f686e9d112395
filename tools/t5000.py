import sys; sys.path.insert(0,'/root/repo')
import __graft_entry__ as ge, numpy as np
pkg=ge.load_package(); orc=ge.load_oracle()
pkg.lib().bphf_set_build_mode(0,3000)
for n in (5000, 9000):
    keys=orc.keygen(n, seed=n+3, kind=0)
    try:
        m=pkg.Mphf.new_parallel(1.7, keys, None); print(n, "ok", m.build_stats()["bucketed_levels"])
    except Exception as e: print(n, "ERR", e)
