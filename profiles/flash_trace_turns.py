import sys
sys.argv=['x','64','1']
exec(open('profiles/flash_trace.py').read().split("print(f\"hd=")[0])
import numpy as np
viol=0
for j in range(2,31):
    a = rel[1,j,3]-rel[0,j,4]   # tile1 turn(j) - tile0 P_issued(j)  (should be >= 0)
    b = rel[0,j+1,3]-rel[1,j,4] # tile0 turn(j+1) - tile1 P_issued(j)
    print(j, int(a), int(b))
