import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import pyoracle
from measure_ia_b200 import MeasureIABox
from measure_ia_b200.synthetic import uniform_box
n=int(sys.argv[1]) if len(sys.argv)>1 else 3000
jk=int(sys.argv[2]) if len(sys.argv)>2 else 0
d=uniform_box(n,205.0,seed=1)
for kern in ('general','tiled'):
    b=MeasureIABox(d,None,boxsize=205.0,num_bins_r=10,num_bins_pi=8); b.kernel=kern
    b.measure_xi_w('a','both',jk,temp_file_path=False)
    r=b.last_result
    print(kern, b.last_stats['tested'], b.last_stats['binned'], r['count'].sum())
    if kern=='general': ref=r
    else:
        print('count diff:\n', r['count']-ref['count'])
        print('spd rel diff max', np.abs(r['SpD_raw']-ref['SpD_raw']).max(), 'scd', np.abs(r['ScD_raw']-ref['ScD_raw']).max(), np.abs(ref['ScD_raw']).max())
