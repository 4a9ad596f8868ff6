#!/usr/bin/env python
"""Symbolic comparison of straight-line fp32 SASS: which product of every  a*b +- c*d  is rounded and which is fused is the
compiler's choice and depends on the surrounding code, so "the same expression tree" in the source does not guarantee the
same bits.  This tool evaluates the float instructions of a kernel listing (cuobjdump -sass) symbolically along the fast
paths (the branches around the slow-path CALLs of divisions are taken), with x*0 = 0 and a*b+0 = round(a*b) folded, and
compares the expression of the 2-D covariance determinant -- which every conic entry, hence alpha and T, depends on -- of the
reference's preprocessCUDA (oracle/_ref/libgsref.so) with both instantiations of preprocess_forward_kernel.

    python tools/sass_symbolic.py          # prints IDENTICAL / DIFFERENT (+ the smallest differing subtrees) per kernel

Round 2 found one difference this way (x*z +- r*y in the rotation matrix: the reference rounds x*z, the product build
rounded r*y; 14 % of the conic entries differed by up to 1e-5 relative) -- now pinned with intrinsics in preprocess.cu.
Test infrastructure / diagnosis only; needs cuobjdump and both libraries."""
import itertools
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Symbolic straight-line evaluation of float SASS (fast paths only): builds expression strings per register.
def parse(path):
    ins=[]
    for ln in open(path):
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', ln)
        if m: ins.append((int(m.group(1),16), m.group(2).strip()))
    return ins
def run(path, start=None, stop=None, names=None, verbose=False):
    regs={}
    names=names or {}
    loads=[]
    def val(tok):
        tok=tok.strip()
        neg=False; ab=False
        if tok.startswith('-'): neg=True; tok=tok[1:]
        if tok.startswith('|') and tok.endswith('|'): ab=True; tok=tok[1:-1]
        tok=tok.replace('.reuse','')
        if tok=='RZ' or tok=='URZ': v='0'
        elif re.match(r'^U?R\d+$',tok): v=regs.get(tok, tok+'?')
        elif tok.startswith('c['): v=names.get(tok,tok)
        else:
            try: v=repr(float(tok))
            except: v=tok
        if ab: v='abs(%s)'%v
        if neg: v='neg(%s)'%v if v!='0' else '0'
        return v
    out=[]
    allins=parse(path)
    skip_to=None
    for idx,(addr,txt) in enumerate(allins):
        if start is not None and addr<start: continue
        if stop is not None and addr>=stop: break
        if skip_to is not None:
            if addr<skip_to: continue
            skip_to=None
        mb=re.match(r'^@!?U?P\d+\s+BRA\s+(?:U,)?\s*(0x[0-9a-f]+)',txt)
        if mb:
            tgt=int(mb.group(1),16)
            if tgt>addr and tgt-addr<=0x90 and any('CALL' in t for a,t in allins[idx+1:] if a<tgt):
                skip_to=tgt      # fast path: the branch around a slow-path CALL is taken
            continue
        pred=None
        m=re.match(r'^(@!?U?P\d+)\s+(.*)',txt)
        if m: pred,txt=m.group(1),m.group(2)
        parts=txt.split(None,1)
        op=parts[0]; args=[a.strip() for a in parts[1].split(',')] if len(parts)>1 else []
        base=op.split('.')[0]
        if pred and base in ('BRA','CALL','EXIT','MOV','IMAD','MUFU'):   # slow paths / predicated moves: ignore
            if base in ('MUFU',):
                pass
            else:
                continue
        d=args[0].replace('.reuse','') if args else None
        def mul_(a,b):
            if a=='0' or b=='0': return '0'          # finite operands: x * 0 = 0 exactly
            return 'mul(%s,%s)'%tuple(sorted([a,b]))
        if base=='FMUL':
            regs[d]=mul_(val(args[1]),val(args[2]))
        elif base=='FADD':
            a,b=val(args[1]),val(args[2])
            regs[d]=b if a=='0' else (a if b=='0' else 'add(%s,%s)'%tuple(sorted([a,b])))
        elif base=='FFMA':
            a,b,c=val(args[1]),val(args[2]),val(args[3])
            if a=='0' or b=='0': regs[d]=c                      # 0 * x + c = c
            elif c=='0': regs[d]=mul_(a,b)                       # a * b + 0 = round(a * b)
            else:
                a,b=sorted([a,b])
                regs[d]='fma(%s,%s,%s)'%(a,b,c)
        elif base=='FMNMX':
            regs[d]='%s(%s,%s)'%('min' if args[3].strip()=='PT' else 'max', val(args[1]),val(args[2]))
        elif base=='MUFU':
            regs[d]='%s(%s)'%(op.split('.')[1].lower(), val(args[1]))
        elif base in ('MOV',):
            regs[d]=val(args[1])
        elif base=='IMAD' and ('MOV' in op or op=='IMAD.U32'):
            regs[d]=val(args[3]) if len(args)>3 else val(args[1])
        elif base in ('LDG','LDS','LDC','LDCU','LD'):
            key='%x:%s %s'%(addr,op,','.join(args[1:]))
            nm=names.get('%x'%addr)
            wide = 4 if '.128' in op else (2 if '.64' in op else 1)
            m2=re.match(r'^(U?R)(\d+)$',d)
            for k in range(wide):
                r='%s%d'%(m2.group(1),int(m2.group(2))+k)
                regs[r]=(nm[k] if isinstance(nm,(list,tuple)) else (nm if nm and wide==1 else 'L%x_%d'%(addr,k)))
            loads.append(key)
        elif base in ('F2F','I2F','F2I','DMUL','DADD','DFMA'):
            regs[d]='%s(%s)'%(op,','.join(val(a) for a in args[1:]))
        else:
            if d and re.match(r'^U?R\d+$',d) and base not in ('STG','STS','ST','BSSY','BSYNC','FCHK','FSETP','ISETP','BRA','CALL'):
                regs[d]='%s@%x'%(base,addr)
        if verbose: out.append((addr,txt,regs.get(d) if d else None))
    return regs, loads, out


def auto_names(path):
    ins=parse(path)
    names={}
    stop=None
    for a,t in ins:
        if '0.30000001192' in t: stop=a; break
    scal=[]; ldg128=[]
    for a,t in ins:
        if a>=stop: break
        m=re.match(r'LDS\.128 R\d+, \[U?R\d+\+0x(24[0-7][0-9a-f])\]',t)
        if m:
            off=int(m.group(1),16)
            if off<0x2440: k=(off-0x2400)//4; names['%x'%a]=['vm%d'%(k+i) for i in range(4)]
            else: k=(off-0x2440)//4; names['%x'%a]=['pj%d'%(k+i) for i in range(4)]
            continue
        m=re.match(r'LDS R\d+, \[R\d+(\+0x([48]))?\]',t)
        if m: scal.append((a,int(m.group(2),16) if m.group(2) else 0)); continue
        if t.startswith('LDG.E.128'): ldg128.append(a)
        if re.match(r'LDCU UR\d+, c\[0x0\]\[0x3f8\]',t): names['%x'%a]='mod'
        for off,nm in (('0x3fc','tanx'),('0x400','tany'),('0x404','fx'),('0x408','fy')):
            if re.match(r'LDCU? U?R\d+, c\[0x0\]\[%s\]'%off,t): names['%x'%a]=nm
    comp={0:'x',4:'y',8:'z'}
    for i,(a,off) in enumerate(scal[:6]):
        names['%x'%a]=('p' if i<3 else 's')+comp[off]
    names['%x'%ldg128[-1]]=['qr','qx','qy','qz']
    return names, stop


def parse_expr(s):
    # returns nested tuple
    pos=0
    def p():
        nonlocal pos
        m=re.match(r'[A-Za-z_0-9\.\[\]\?@x-]+',s[pos:])
        name=m.group(0); pos+=len(name)
        if pos<len(s) and s[pos]=='(':
            pos+=1; args=[]
            while True:
                args.append(p())
                if s[pos]==',': pos+=1
                elif s[pos]==')': pos+=1; break
            return (name,)+tuple(args)
        return name
    return p()
def size(t): return 1 if isinstance(t,str) else 1+sum(size(a) for a in t[1:])
def show(t,d=0,maxd=3):
    if isinstance(t,str): return t
    if d>=maxd: return '<%s:%d>'%(t[0],size(t))
    return '%s(%s)'%(t[0],','.join(show(a,d+1,maxd) for a in t[1:]))
def diff(a,b,path='',out=None):
    if out is None: out=[]
    if a==b: return out
    if isinstance(a,str) or isinstance(b,str) or a[0]!=b[0] or len(a)!=len(b):
        out.append((path,a,b)); return out
    xs,ys=list(a[1:]),list(b[1:])
    fixed=[]
    if a[0]=='fma':
        if xs[2]!=ys[2]: fixed.append((xs[2],ys[2]))
        xs,ys=xs[:2],ys[:2]
    if a[0] in ('mul','add','fma'):
        for x in list(xs):
            if x in ys: xs.remove(x); ys.remove(x)
        best=None
        for perm in itertools.permutations(ys):
            o=[]
            for x,y in zip(xs,perm): diff(x,y,path+'/'+a[0],o)
            if best is None or len(o)<len(best): best=o
        out.extend(best or [])
    else:
        for x,y in zip(xs,ys): diff(x,y,path+'/'+a[0],out)
    for x,y in fixed: diff(x,y,path+'/'+a[0]+'.c',out)
    return out
def cmp(label,rs,os_):
    a,b=parse_expr(rs),parse_expr(os_)
    d=diff(a,b)
    print('==',label,'sizes',size(a),size(b),'EQUAL' if not d else '%d diffs'%len(d))
    seen=set()
    for path,x,y in d[:200]:
        k=(show(x,0,4),show(y,0,4))
        if k in seen: continue
        seen.add(k)
        print('  at',path[-60:]); print('    ref:',k[0]); print('    our:',k[1])
        if len(seen)>6: break


REF_NAMES = {'3a0': 'px', '3d0': 'py', '3f0': 'pz', '540': 'py', '560': 'px', '570': 'pz', '960': 'qx', '970': 'qr', '980': 'qz', '9a0': 'qy',
             '9b0': 'sy', '9c0': 'sx', '9d0': 'sz', 'f00': 'vm6', 'f10': 'vm2', 'f20': 'vm10', 'f30': 'vm14', 'f40': 'vm4', 'f50': 'vm0',
             'f60': 'vm8', 'f70': 'vm12', 'f80': 'vm5', 'f90': 'vm1', 'fa0': 'vm9', 'fb0': 'vm13', '930': 'mod', '1170': ['tanx', 'tany'],
             '1270': 'fx', '13d0': 'fx', '14f0': 'fy', '1610': 'fy', '1360': 'fx', '15a0': 'fy'}      # load address -> name (nvcc 12.9 build)


def dump(lib, pattern, out):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()
    keep, on = [], False
    for ln in txt:
        if "Function :" in ln:
            if on:
                break
            on = re.search(pattern, ln) is not None
        if on:
            keep.append(ln)
    open(out, "w").write("\n".join(keep) + "\n")
    return out


def det_of(path, names):
    stop = None
    for a, t in parse(path):
        if "0.30000001192" in t:
            stop = a
            break
    for a, t in parse(path):
        if a > stop and re.match(r"FFMA R\d+, R\d+(\.reuse)?, R\d+(\.reuse)?, -R\d+", t):
            regs, _, _ = run(path, 0, a + 0x10, names)
            return regs[t.split()[1].rstrip(",")]
    raise RuntimeError("determinant not found in " + path)


def subtrees(t, acc):
    if isinstance(t, str):
        return
    acc.add(t)
    for c in t[1:]:
        subtrees(c, acc)


if __name__ == "__main__":
    tmp = os.environ.get("TMPDIR", "/tmp")
    ref = dump(os.path.join(ROOT, "oracle", "_ref", "libgsref.so"), r"_Z14preprocessCUDAILi3EEviiiPKfPKN3glm", os.path.join(tmp, "ref_pre.sass"))
    ours = os.path.join(ROOT, "gs-slam-analytica_jacobian_b200", "diff_gaussian_rasterization", "libgsr_b200.so")
    r = det_of(ref, REF_NAMES)
    rc = 0
    for inst in ("ILb0", "ILb1"):
        f = dump(ours, "preprocess_forward_kernel" + inst, os.path.join(tmp, "our_pre_%s.sass" % inst))
        names, _ = auto_names(f)
        o = det_of(f, names)
        print("preprocess_forward_kernel<%s>: cov2D determinant expression %s" % (inst[3:], "IDENTICAL to the reference's" if o == r else "DIFFERENT"))
        if o != r:
            rc = 1
            A, B = parse_expr(r), parse_expr(o)
            sa, sb = set(), set()
            subtrees(A, sa); subtrees(B, sb)
            for t in sorted(sa - sb, key=size)[:5]:
                print("   reference only (%d nodes): %s" % (size(t), show(t, 0, 5)))
            for t in sorted(sb - sa, key=size)[:5]:
                print("   this build only (%d nodes): %s" % (size(t), show(t, 0, 5)))
    sys.exit(rc)
