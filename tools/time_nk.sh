# N-k batches on the ringed 1804-node grid (tools/time_n1.py with K branches per scenario): shared pattern against one by one
for K in 1 2 3; do N_SCN=1000 K=$K NODES=1500 PGMB_DEBUG_N1=1 python tools/time_n1.py 2>&1; done
N_SCN=1000 K=2 NODES=1500 ORACLE=0 PGMB_OUTAGE_SLOTS=1 python tools/time_n1.py
N_SCN=1000 K=2 NODES=1500 ASYM=1 python tools/time_n1.py
N_SCN=200 K=2 NODES=50000 PGMB_DEBUG_N1=1 python tools/time_n1.py
