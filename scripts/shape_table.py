#!/usr/bin/env python
"""Shape-mapped table of one denoising step from an ncu launch list (gpu__time_duration.sum per launch): every conv /
GroupNorm / attention launch of the CIFAR-10 UNet is labelled with its layer shape (the launch order is the block
order of unet.py:286-322) and reported with its TFLOP/s (algorithmic FLOPs) or TB/s (minimum bytes: one read of the
input + one 16-bit write).  usage: shape_table.py launches.csv [index of the step in the list, default last] [rows]"""
import csv,sys,collections
path=sys.argv[1]
lines=[l for l in open(path) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
idx=[i for i,r in enumerate(rows) if 'begin_step' in r['Kernel Name']]
s=rows[idx[int(sys.argv[2]) if len(sys.argv)>2 else -1]:]
# build labels
labels=[]
L=3; nrb=3; attn=[0,1,1]; hid=256
res=32
# CFG: in_conv, norm1 and conv1 of block 0 run once per sample = on half of the rows (plan.cu: cfg_prefix_shareable)
SHARED=' [shared by the CFG pair: R/2 rows]'
def add_res(name,cin,cout,rs,level,shared=False):
    global res
    ro = res//2 if rs=='down' else res*2 if rs=='up' else res
    f = 0.5 if shared else 1
    sfx = SHARED if shared else ''
    labels.append(('GN', f'norm1 C{cin}@{res}'+(' '+rs if rs else '')+(' +raw' if cin!=cout else '')+sfx, cin*res*res*f))
    labels.append(('CONV', f'conv1 3x3 {cin}->{cout}@{ro}'+(' subpixel' if rs=='up' else '')+sfx, 2*ro*ro*cout*9*cin*f))
    labels.append(('GN', f'norm2 C{cout}@{ro} in16', cout*ro*ro))
    labels.append(('CONV', f'conv2 3x3 {cout}->{cout}@{ro}'+(f'+skip1x1 {cin}' if cin!=cout else ''), 2*ro*ro*cout*(9*cout+(cin if cin!=cout else 0))))
    res=ro
    if attn[level]: add_attn(cout)
def add_attn(c):
    labels.append(('GN', f'attn.norm C{c}@{res}', c*res*res))
    labels.append(('CONV', f'proj_in 1x1 {c}->{3*c}@{res}', 2*res*res*c*3*c))
    labels.append(('ATTN', f'attn N={res*res}', 4*(res*res)**2*c))
    labels.append(('CONV', f'proj_out 1x1 {c}->{c}@{res}', 2*res*res*c*c))
import os
share = os.environ.get('VDT_NO_CFG_SHARE') is None
labels.append(('CONV','in_conv'+(SHARED if share else ''),2*32*32*256*27*(0.5 if share else 1)))
for i in range(L):
    for j in range(nrb): add_res('d',256,256,None,i,shared=(share and i==0 and j==0))
    if i!=L-1: add_res('d',256,256,'down',i)
add_res('m',256,256,None,0) if False else None
# middle: Res, Attn, Res (no level attn flag use)
def add_res_na(cin,cout):
    global res
    labels.append(('GN', f'norm1 C{cin}@{res}', cin*res*res))
    labels.append(('CONV', f'conv1 3x3 {cin}->{cout}@{res}', 2*res*res*cout*9*cin))
    labels.append(('GN', f'norm2 C{cout}@{res} in16', cout*res*res))
    labels.append(('CONV', f'conv2 3x3 {cout}->{cout}@{res}', 2*res*res*cout*9*cout))
add_res_na(256,256); add_attn(256); add_res_na(256,256)
for i in reversed(range(L)):
    for j in range(nrb+1): add_res('u',512,256,None,i)
    if i!=0: add_res('u',256,256,'up',i)
labels.append(('GN','out.norm C256@32',256*1024))
labels.append(('CONV','out_conv 3x3 256->3',2*1024*3*9*256))
# zip with launches
k=0
agg=collections.OrderedDict()
for r in s:
    n=r['Kernel Name']
    t=float(r['Metric Value'])/1e3
    if 'conv_gemm' in n: kind='CONV'
    elif 'groupnorm_kernel' in n: kind='GN'
    elif 'attention_kernel' in n: kind='ATTN'
    else: continue
    if k>=len(labels): break
    lk,name,work=labels[k]
    assert lk==kind,(k,lk,kind,name)
    a=agg.setdefault((kind,name),[0,0.0,work]); a[0]+=1; a[1]+=t
    k+=1
print('matched',k,'of',len(labels))
R=int(sys.argv[3]) if len(sys.argv)>3 else 1024
tot=collections.defaultdict(float)
for (kind,name),(n,t,work) in agg.items():
    tot[kind]+=t
    if kind=='GN':
        in16='in16' in name
        b=work*R*((2 if in16 else 4)+2)
        extra=''
        print(f'{kind:5s} {name:40.40s} n={n:3d} avg_us={t/n:8.1f} total_ms={t/1e3:7.3f}  min-bytes {b/1e6:8.1f} MB -> {b/(t/n*1e-6)/1e12:5.2f} TB/s')
    else:
        print(f'{kind:5s} {name:40s} n={n:3d} avg_us={t/n:8.1f} total_ms={t/1e3:7.3f}  {work*R/(t/n*1e-6)/1e12:7.1f} TFLOP/s')
print(dict(tot))
