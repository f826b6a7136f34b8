"""Development aid: for every short loop of shared-memory loads and FP64 adds in an object file, how many LDS are issued before the
first DADD (1 = every add waits for the load right before it; the stand-alone product kernel has 11).  usage: sass_sched.py file.o"""
import re,sys,subprocess
obj=sys.argv[1]
txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
ins=[]
for l in txt.splitlines():
    if 'Function :' in l: ins.append((-1,l.strip())); continue
    m=re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1),16), m.group(2).strip()))
cur=None; idx={}
for i,(a,t) in enumerate(ins):
    if a==-1: cur=re.sub(r'.*?(k_\w+?|product_pass)I(Lb\d)E.*',r'\1<\2>',t); idx={}; continue
    idx[a]=i
    if 'BRA' in t:
        m2=re.search(r'0x([0-9a-f]+)', t)
        if m2:
            tgt=int(m2.group(1),16)
            if tgt<=a and tgt in idx and (i-idx[tgt])<80:
                body=[x[1].split()[0] if not x[1].startswith('@') else x[1].split()[1] for x in ins[idx[tgt]:i+1]]
                if sum(x.startswith('LDS') for x in body)>=20:
                    # leading LDS run length before first DADD
                    k=0
                    for x in body:
                        if x.startswith('DADD'): break
                        if x.startswith('LDS'): k+=1
                    print(cur[:50], hex(tgt), 'LDS before first DADD:',k)
