// CPU emulation (lane by lane, shuffle by shuffle) of swipe_prof_kernel<R,true> + walk_kernel, checked against the oracle.
// Validates the design arguments the CUDA kernel relies on: -128 profile halo instead of validity tests, target clamp to
// the delimiters, dead-row select, wavefront-major nibble trace.  Built and run by tests/test_swipe_emulation.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <random>
#include "../include/dmnd_b200.h"
using namespace std;
static int8_t S[1024];
static inline int sat8(int v){ return v>127?127:(v<-128?-128:v); }
template<int R> int run(const int8_t* q, const int8_t* cb, int qlen, const int8_t* t, int tlen, int d_begin, int d_end, dmnd_dp_result& res) {
	const int B = d_end - d_begin; const int i1 = max(d_end - 1, 0), j0 = i1 - (d_end - 1);
	const int cols = min(qlen - 1 - d_begin, tlen - 1) + 1 - j0;
	memset(&res, 0, sizeof res);
	if (B <= 0 || cols <= 0) return 0;
	const int HALO = 16*R, W = qlen + 32*R + 4;
	vector<int8_t> prof((size_t)27*W, -128);
	bool ok=true;
	for (int idx=0; idx<W; ++idx){ int i = idx-HALO; if (i>=0 && i<qlen) for(int a=0;a<26;++a){ int v=S[a*32+(q[i]&31)]+cb[i]; if(v>127||v<-127) ok=false; prof[(size_t)a*W+idx]=(int8_t)sat8(v);} }
	if(!ok) return -1;
	const int nsteps = 2 * (cols - 1) + B, nmacro = (nsteps + 1) >> 1;
	const int ibase = j0 + d_begin;
	const int m_lo = max(0, -ibase - 16*R), m_hi = min(nmacro, qlen - ibase);
	vector<uint8_t> tr((size_t)max(nmacro,1)*16*R, 0xEE);
	static int H[32][R], E[32][R], F[32][R], bestv[32][R], bestc[32][R]; static int trow[32][R/2];
	memset(H,0,sizeof H); memset(E,0,sizeof E); memset(F,0,sizeof F); memset(bestv,0,sizeof bestv); memset(bestc,0,sizeof bestc);
	const int go=12, ge=1;
	auto tl = [&](int j){ int jj = min(max(j,-1), tlen); int l = t[jj]&31; return min(l,26)*W; };
	for (int lane=0;lane<32;++lane) for(int u=0;u<R/2;++u) trow[lane][u] = tl(j0 + m_lo - lane*R/2 - u);
	for (int m = m_lo; m < m_hi; ++m) {
		int fup[32], edn[32]; uint32_t pk[32][4]; memset(pk,0,sizeof pk);
		for (int lane=0;lane<32;++lane) fup[lane] = lane ? F[lane-1][R-1] : 0;
		for (int lane=0;lane<32;++lane){ const int kb = min(R-1, B-1-lane*R); const int I0 = ibase + m + lane*R/2 + HALO;
			for (int k=0;k<R;k+=2){ const int u=k>>1; const int c = m - lane*R/2 - u;
				const int sc = prof[(size_t)trow[lane][u] + I0 + u];
				const int e_in=E[lane][k+1], f_in = k>0?F[lane][k-1]:fup[lane];
				const int hd=H[lane][k]+sc; const int h= k<=kb ? max(max(hd,e_in),max(f_in,0)) : 0; const int open=max(h-go,0);
				const int e=max(max(e_in-ge,0),open), f=max(max(f_in-ge,0),open);
				int b0=(f_in>=hd)&(f_in>=e_in), b1=(e_in>=hd)&(e_in>=f_in), b2= open>=f_in-ge, b3 = open>=e_in-ge;
				pk[lane][k>>3] |= (uint32_t)(b0|(b1<<1)|(b2<<2)|(b3<<3)) << ((k&7)*4);
				if (h>bestv[lane][k]){bestv[lane][k]=h;bestc[lane][k]=c;}
				H[lane][k]=h;E[lane][k]=e;F[lane][k]=f; } }
		for (int lane=0;lane<32;++lane) edn[lane] = lane<31 ? E[lane+1][0] : 0;
		for (int lane=0;lane<32;++lane){ const int kb = min(R-1, B-1-lane*R); const int I0 = ibase + m + lane*R/2 + HALO;
			for (int k=1;k<R;k+=2){ const int u=k>>1; const int c = m - lane*R/2 - u;
				const int sc = prof[(size_t)trow[lane][u] + I0 + u + 1];
				const int e_in = k+1<R?E[lane][k+1]:edn[lane], f_in=F[lane][k-1];
				const int hd=H[lane][k]+sc; const int h= k<=kb ? max(max(hd,e_in),max(f_in,0)) : 0; const int open=max(h-go,0);
				const int e=max(max(e_in-ge,0),open), f=max(max(f_in-ge,0),open);
				int b0=(f_in>=hd)&(f_in>=e_in), b1=(e_in>=hd)&(e_in>=f_in), b2= open>=f_in-ge, b3 = open>=e_in-ge;
				pk[lane][k>>3] |= (uint32_t)(b0|(b1<<1)|(b2<<2)|(b3<<3)) << ((k&7)*4);
				if (h>bestv[lane][k]){bestv[lane][k]=h;bestc[lane][k]=c;}
				H[lane][k]=h;E[lane][k]=e;F[lane][k]=f; } }
		for (int lane=0;lane<32;++lane){ for(int x=0;x<R/2;++x) tr[((size_t)m*32+lane)*(R/2)+x] = (uint8_t)(pk[lane][x>>2] >> ((x&3)*8));
			for (int u=R/2-1;u>0;--u) trow[lane][u]=trow[lane][u-1]; trow[lane][0] = tl(j0 + (m+1) - lane*R/2); }
	}
	int bv=0,bc=0,br=0;
	for (int lane=0;lane<32;++lane) for (int k=0;k<R;++k){ int r=lane*R+k; if (r>=B) continue; if (bestv[lane][k] > bv || (bestv[lane][k]==bv && bv>0 && (bestc[lane][k] < bc || (bestc[lane][k]==bc && r>br)))) {bv=bestv[lane][k];bc=bestc[lane][k];br=r;} }
	res.score = bv;
	auto nibf=[&](int c,int r){ int m=c+(r>>1), lane=r/R, k=r-lane*R; uint8_t b=tr[((size_t)m*32+lane)*(R/2)+(k>>1)]; return (b>>((k&1)*4))&15; };
	if (bv > 0) {
		int c = bc, r = br, i = j0 + d_begin + c + r, j = j0 + c; res.q_end=i+1; res.t_end=j+1; int sc=0; bool bad=false;
		while (i>=0 && j>=0 && sc<bv) {
			if (c<0||r<0||r>=B){bad=true;break;}
			int nib = nibf(c,r);
			if ((nib&3)==0){ int ql=q[i]&31, sl=t[j]&31; sc += S[(ql<<5)|sl] + cb[i]; if(ql==sl)++res.identities; else ++res.mismatches; ++res.length; --i;--j;--c; }
			else if (nib&1){ int l=0; do{++l;--i;--r;} while (r>=0 && (nibf(c,r)&4)==0 && i>0); if(r<0){bad=true;break;} ++res.gap_openings; res.length+=l; sc -= 11 + l; }
			else { int l=0; do{++l;--j;--c;++r;} while (c>=0 && r<B && (nibf(c,r)&8)==0 && j>0); if(c<0||r>=B){bad=true;break;} ++res.gap_openings; res.length+=l; sc -= 11+l; }
		}
		if (bad || sc != bv) res.status = 2;
		res.q_begin=i+1; res.t_begin=j+1;
	}
	return 0;
}
int main(int argc, char** argv) {
	dmnd_search_opts o; dmnd_search_opts_default(&o); dmnd_params p; dmnd_params_init(&o,&p); memcpy(S,p.score,1024);
	dmnd_ctx* ctx; dmnd_create(0,&p,&ctx);
	mt19937 rng(atoi(argc>1?argv[1]:"1")); int maxq = argc>2?atoi(argv[2]):300, maxw = argc>3?atoi(argv[3]):200, iters=argc>4?atoi(argv[4]):3000;
	int fails=0, pos=0;
	for (int it=0; it<iters; ++it) {
		int qlen = 5 + rng()%maxq, tlen = 5 + rng()%(maxq*4/3);
		// neighbours before/after both sequences are REAL letters (other sequences), to catch reads across delimiters
		vector<int8_t> qb(256+60+1+qlen+1+60+1+256,31), tb(256+60+1+tlen+1+60+1+256,31);
		for(int i=0;i<60;++i){ qb[256+i]=rng()%20; tb[256+i]=rng()%20; qb[256+61+qlen+1+i]=rng()%20; tb[256+61+tlen+1+i]=rng()%20; }
		const int qo=256+61, to=256+61;
		for(int i=0;i<tlen;++i) tb[to+i]=rng()%20;
		int off = rng()%max(1,tlen-qlen+1);
		for(int i=0;i<qlen;++i) qb[qo+i]= (rng()%100<75 && off+i<tlen) ? tb[to+off+i] : rng()%20;
		if (rng()%5==0) qb[qo + rng()%qlen] = 23;
		int64_t qlim[4]={256,256+61,256+61+qlen+1,256+61+qlen+1+61}, tlim[4]={256,256+61,256+61+tlen+1,256+61+tlen+1+61};
		int lo=-(tlen-1), hi=qlen; int c = -off + (int)(rng()%41) - 20; int w = 1 + rng()%maxw; int kind=rng()%10; int d0,d1;
		if (kind<7){ d0=max(lo,c-w/2); d1=min(hi,d0+w);} else if(kind==7){ d0=lo; d1=min(hi,lo+w);} else if(kind==8){ d1=hi; d0=max(lo,hi-w);} else { d0=lo+rng()%(hi-lo); d1=d0+1; }
		if(d1<=d0) d1=d0+1;
		dmnd_block *bq,*bt; dmnd_block_upload(ctx,qb.data(),qb.size(),qlim,3,&bq); dmnd_block_upload(ctx,tb.data(),tb.size(),tlim,3,&bt);
		vector<int8_t> bias(qb.size(),0); for(auto&x:bias) x = (int8_t)((int)(rng()%5)-3); dmnd_block_set_bias(ctx,bq,bias.data(),bias.size());
		dmnd_dp_problem pr{1,1,d0,d1}; dmnd_dp_result ro, re;
		if (dmnd_banded_swipe(ctx,bq,bt,&pr,1,1,&ro,nullptr,0)) { printf("oracle err %s\n", dmnd_last_error()); }
		int B=d1-d0, rc;
		const int8_t* Q=qb.data()+qo; const int8_t* CB=bias.data()+qo; const int8_t* T=tb.data()+to;
		if (B<=64) rc=run<2>(Q,CB,qlen,T,tlen,d0,d1,re); else if (B<=128) rc=run<4>(Q,CB,qlen,T,tlen,d0,d1,re); else if (B<=256) rc=run<8>(Q,CB,qlen,T,tlen,d0,d1,re); else if (B<=512) rc=run<16>(Q,CB,qlen,T,tlen,d0,d1,re); else rc=run<32>(Q,CB,qlen,T,tlen,d0,d1,re);
		if (rc) { printf("profile overflow\n"); continue; }
		if (ro.score>0) ++pos;
		bool same = ro.score==re.score && ro.q_begin==re.q_begin && ro.q_end==re.q_end && ro.t_begin==re.t_begin && ro.t_end==re.t_end && ro.identities==re.identities && ro.mismatches==re.mismatches && ro.gap_openings==re.gap_openings && ro.length==re.length && re.status==0;
		if(!same){ ++fails; if(fails<6) printf("MISMATCH it=%d qlen=%d tlen=%d d0=%d d1=%d oracle score=%d q[%d,%d) t[%d,%d) | emu score=%d q[%d,%d) t[%d,%d) status=%d\n",it,qlen,tlen,d0,d1,ro.score,ro.q_begin,ro.q_end,ro.t_begin,ro.t_end,re.score,re.q_begin,re.q_end,re.t_begin,re.t_end,re.status); }
		dmnd_block_free(ctx,bq); dmnd_block_free(ctx,bt);
	}
	printf("fails=%d positives=%d\n",fails,pos);
}
