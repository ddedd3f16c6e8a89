"""Pure-Python restatement of the CABAC bit count of a TU's quantised levels -- TEST INFRASTRUCTURE ONLY.

Follows kvz_encode_coeff_nxn in only_count mode (reference: src/strategies/generic/encode_coding_tree-generic.c:40-290),
kvz_encode_last_significant_xy (src/encode_coding_tree.c:63-115) and kvz_cabac_write_coeff_remain (src/cabac.c:275-301),
i.e. the CABAC branch of kvz_get_coeff_cost (src/rdo.c:223-264, 291-330), without context adaptation.
Pinned against the compiled reference by tests/test_rdoq.py::test_python_coeff_cost_port_vs_reference (CPU).
Nothing in kvazaar_b200/ imports this file.
"""
import numpy as np

MPS=[32768,30426,28306,26378,24617,23005,21523,20159,18899,17734,16653,15650,14717,13849,13038,12282,11575,10914,10294,9714,9169,8658,8178,7727,7303,6903,6527,6173,5840,5525,5228,4948,4684,4435,4199,3977,3767,3568,3380,3202,3034,2876,2725,2583,2448,2321,2200,2086,1978,1875,1778,1686,1599,1517,1439,1364,1294,1228,1165,1105,1048,994,943,895]
LPS=[32768,35232,37696,40159,42623,45087,47551,50015,52479,54942,57406,59870,62334,64798,67262,69725,72189,74653,77117,79581,82044,84508,86972,89436,91900,94363,96827,99291,101755,104219,106683,109146,111610,114074,116538,119002,121465,123929,126393,128857,131321,133785,136248,138712,141176,143640,146104,148568,151031,153495,155959,158423,160887,163351,165814,168278,170742,173207,175669,178134,180598,183061,185525,187989]
EB=[ (LPS[i>>1] if i&1 else MPS[i>>1]) for i in range(128)]
names=[("sao_merge_flag_model",1),("sao_type_idx_model",1),("split_flag_model",3),("intra_mode_model",1),("chroma_pred_model",2),("inter_dir",5),("trans_subdiv_model",3),("qt_cbf_model_luma",4),("qt_cbf_model_chroma",4),("cu_qp_delta_abs",4),("part_size_model",4),("cu_sig_coeff_group_model",4),("cu_sig_model_luma",27),("cu_sig_model_chroma",15),("cu_ctx_last_y_luma",15),("cu_ctx_last_y_chroma",15),("cu_ctx_last_x_luma",15),("cu_ctx_last_x_chroma",15),("cu_one_model_luma",16),("cu_one_model_chroma",8),("cu_abs_model_luma",4),("cu_abs_model_chroma",2),("cu_pred_mode_model",1),("cu_skip_flag_model",3),("cu_merge_idx_ext_model",1),("cu_merge_flag_ext_model",1),("cu_transquant_bypass",1),("cu_mvd_model",2),("cu_ref_pic_model",2),("mvp_idx_model",2),("cu_qt_root_cbf_model",1),("transform_skip_model_luma",1),("transform_skip_model_chroma",1)]
OFF={};o=0
for nme,c in names: OFF[nme]=o;o+=c
def last_group(x):
    if x<4: return x
    l=x.bit_length()-1
    return 2*l+((x>>(l-1))&1)
def sig_ctx(pattern,scan,px,py,log2n,typ):
    if px+py==0: return 0
    if log2n==2: return [0,1,4,5,2,3,4,5,6,6,8,8,7,7,8,8][4*py+px]
    offset=(9 if scan==0 else 15) if log2n==3 else (21 if typ==0 else 12)
    sx,sy=px&3,py&3
    if pattern==0: cnt=(2 if sx+sy==0 else 1) if sx+sy<=2 else 0
    elif pattern==1: cnt=(2 if sy==0 else 1) if sy<=1 else 0
    elif pattern==2: cnt=(2 if sx==0 else 1) if sx<=1 else 0
    else: cnt=2
    return (3 if (typ==0 and ((px>>2)+(py>>2))>0) else 0)+offset+cnt
def remain_bits(symbol,rice):
    if symbol<(3<<rice): return (symbol>>rice)+1+rice
    length=rice; symbol-=3<<rice
    while symbol>=(1<<length): symbol-=1<<length; length+=1
    return 3+length+1-rice+length
def cost(coeff,n,typ,scan_idx,models,scan,signhide,trskip_enable,tr_skip):
    log2n=n.bit_length()-1; side=n>>2; ncg=side*side
    c=coeff.reshape(n,n)
    flags=[0]*ncg
    for g in range(ncg):
        gy,gx=divmod(g,side)
        if c[gy*4:gy*4+4,gx*4:gx*4+4].any(): flags[g]=1
    if not any(flags): return 0.0
    def cg_of_scan(i):
        first=int(scan[i<<4]); return ((first>>log2n)>>2)*side+((first&(n-1))>>2)
    cg_last=ncg-1
    while not flags[cg_of_scan(cg_last)]: cg_last-=1
    scan_last=cg_last*16+15
    flat=coeff.ravel()
    while not flat[scan[scan_last]]: scan_last-=1
    pos_last=int(scan[scan_last])
    bits=0.0
    def binc(acc,off,val):
        return acc+EB[models[off]^val]/32768.0
    if n==4 and trskip_enable: bits=binc(bits,OFF["transform_skip_model_luma"] if typ==0 else OFF["transform_skip_model_chroma"],tr_skip)
    bl=0.0
    lx,ly=pos_last&(n-1),pos_last>>log2n
    if scan_idx==2: lx,ly=ly,lx
    idx=log2n-2
    ctx_offset=0 if typ else idx*3+(idx+1)//4
    shift=idx if typ else (idx+3)//4
    bx=OFF["cu_ctx_last_x_chroma"] if typ else OFF["cu_ctx_last_x_luma"]
    by=OFF["cu_ctx_last_y_chroma"] if typ else OFF["cu_ctx_last_y_luma"]
    gx,gy,gmax=last_group(lx),last_group(ly),last_group(n-1)
    for k in range(gx): bl=binc(bl,bx+ctx_offset+(k>>shift),1)
    if gx<gmax: bl=binc(bl,bx+ctx_offset+(gx>>shift),0)
    for k in range(gy): bl=binc(bl,by+ctx_offset+(k>>shift),1)
    if gy<gmax: bl=binc(bl,by+ctx_offset+(gy>>shift),0)
    if gx>3: bl+=(gx-2)//2
    if gy>3: bl+=(gy-2)//2
    base_cg=OFF["cu_sig_coeff_group_model"]+typ
    base_sig=OFF["cu_sig_model_luma"] if typ==0 else OFF["cu_sig_model_chroma"]
    c1=1; sps=scan_last
    for i in range(cg_last,-1,-1):
        sub=i<<4; ab=[]
        cg_blk=cg_of_scan(i); cgy,cgx=divmod(cg_blk,side)
        last_nz=-1; first_nz=16; rice=0
        if sps==scan_last:
            ab.append(abs(int(flat[pos_last]))); last_nz=sps; first_nz=sps; sps-=1
        right=flags[cgy*side+cgx+1] if cgx<side-1 else 0
        lower=flags[(cgy+1)*side+cgx] if cgy<side-1 else 0
        if i==cg_last or i==0: flags[cg_blk]=1
        else: bits=binc(bits,base_cg+(1 if (right or lower) else 0),flags[cg_blk])
        if flags[cg_blk]:
            pattern=-1 if n==4 else right+(lower<<1)
            while sps>=sub:
                blk=int(scan[sps]); sig=1 if flat[blk]!=0 else 0
                if sps>sub or i==0 or len(ab):
                    bits=binc(bits,base_sig+sig_ctx(pattern,scan_idx,blk&(n-1),blk>>log2n,log2n,typ),sig)
                if sig:
                    ab.append(abs(int(flat[blk])))
                    if last_nz==-1: last_nz=sps
                    first_nz=sps
                sps-=1
        else: sps=sub-1
        if ab:
            nn=len(ab)
            sign_hidden=last_nz-first_nz>=4
            ctx_set=2 if (i>0 and typ==0) else 0
            if c1==0: ctx_set+=1
            c1=1
            base_one=(OFF["cu_one_model_luma"] if typ==0 else OFF["cu_one_model_chroma"])+4*ctx_set
            first_c2=-1
            for k in range(min(nn,8)):
                sym=1 if ab[k]>1 else 0
                bits=binc(bits,base_one+c1,sym)
                if sym:
                    c1=0
                    if first_c2==-1: first_c2=k
                elif 0<c1<3: c1+=1
            if c1==0 and first_c2!=-1:
                bits=binc(bits,(OFF["cu_abs_model_luma"] if typ==0 else OFF["cu_abs_model_chroma"])+ctx_set,1 if ab[first_c2]>2 else 0)
            bits+= nn-1 if (signhide and sign_hidden) else nn
            if c1==0 or nn>8:
                fc2=1
                for k in range(nn):
                    bl_=(2+fc2) if k<8 else 1
                    if ab[k]>=bl_:
                        bits+=remain_bits(ab[k]-bl_,rice)
                        if ab[k]>3*(1<<rice): rice=min(rice+1,4)
                    if ab[k]>=2: fc2=0
    return (0.0+bl)+bits
