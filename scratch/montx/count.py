import re,collections,sys,subprocess
src=sys.argv[1]
cub=src.replace('.cu','.cubin')
subprocess.run(['nvcc','-std=c++17','-O3','-gencode','arch=compute_100a,code=sm_100a','-cubin','-o',cub,src],check=True)
txt=subprocess.run(['cuobjdump','-sass',cub],capture_output=True,text=True).stdout
res=subprocess.run(['cuobjdump','-res-usage',cub],capture_output=True,text=True).stdout
print(re.findall(r'REG:\d+ STACK:\d+',res))
funcs=re.split(r'\n\s*Function : ', txt)
for f in funcs[1:]:
    name=f.split('\n',1)[0]
    ins=[]
    for line in f.split('\n'):
        m=re.match(r'\s+/\*([0-9a-f]{4})\*/\s+((?:@!?U?P[0-9T]+\s+)?)([A-Z0-9_.]+)(.*?);',line)
        if m: ins.append((int(m.group(1),16),m.group(3),m.group(2)+m.group(4)))
    loops=[]
    for a,op,rest in ins:
        if op.startswith('BRA'):
            t=re.search(r'0x([0-9a-f]+)',rest)
            if t and int(t.group(1),16)<a: loops.append((int(t.group(1),16),a))
    # innermost = smallest
    loops.sort(key=lambda l:l[1]-l[0])
    for lo,hi in loops[:2]:
        body=[o for x,o,r in ins if lo<=x<=hi]
        c=collections.Counter(body)
        print(name[:40],'loop %x..%x: %d instrs'%(lo,hi,len(body)), {k:v for k,v in c.most_common(9)})
