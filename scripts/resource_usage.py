"""profiles/r02_resource_usage.txt: registers / stack of every kernel in libcream_b200.so (cuobjdump -res-usage > /tmp/res.txt first)."""
import re,subprocess
txt=open('/tmp/res.txt').read()
out=["# cuobjdump -res-usage cream_b200/libcream_b200.so (sm_100a): registers / stack per kernel (dynamic shared memory is set at launch)"]
fn=None
rows=[]
for line in txt.splitlines():
    m=re.search(r"Function (\S+):",line)
    if m:
        fn=m.group(1); continue
    m=re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)",line)
    if m and fn:
        name=subprocess.run(["c++filt",fn],capture_output=True,text=True).stdout.strip()
        name=re.sub(r"\(anonymous namespace\)::","",name)
        name=re.sub(r"\(.*","",name).replace("void ","").replace("cb::","")
        rows.append((name[:70],int(m.group(1)),int(m.group(2))))
        fn=None
for r in sorted(set(rows)):
    out.append("%-72s regs %3d  stack %3d" % r)
open('profiles/r02_resource_usage.txt','w').write("\n".join(out)+"\n")
print(len(rows))
