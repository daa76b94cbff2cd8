#!/usr/bin/env python3
"""Build a syntax-dumping copy of the reference's hooked HM-16.0 decoder (TEST INFRASTRUCTURE).

Only runs where /root/reference exists (the build container).  It copies
HM_16.0_features/{source,build} to a scratch directory, applies the minimal patch set
listed in SURVEY.md section 0.5 (drop -Werror, drop system("pause"), replace the OpenCV
block of TDecGop.cpp by a binary dump of TDecCu::{bits,MV,CUPU,Pred}), runs the
reference's own linux makefiles on the copy and leaves the binary in oracle/_ref/.
No reference source is copied into this repository.

Dump record (little endian), one per slice/picture:
    int32 poc, int32 nctu, then per CTU (raster order):
        int32 bytes, int32 n_cupu, int32 n_pred, int32 n_mv,
        int16 cupu[n_cupu], int16 pred[n_pred], int16 mv[n_mv]
"""
import os, re, shutil, subprocess, sys

REF = "/root/reference/HM_16.0_features"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "_ref")
WORK = os.environ.get("HM_DUMP_WORK", "/tmp/hm_dump_build")

DUMP = r'''
  /* ---- syntax dump (replaces the reference's OpenCV saliency block) ---- */
  {
    static FILE* s_dump = NULL;
    if (!s_dump) { const char* pth = getenv("HM_SYNTAX_DUMP"); s_dump = fopen(pth ? pth : (getenv("HM_PACKED_FD") ? "/dev/null" : "syntax_dump.bin"), "wb"); }
    int poc = pcSlice->getPOC();
    int nctu = (int)rpcPic->getNumCUsInFrame();
    fwrite(&poc, 4, 1, s_dump); fwrite(&nctu, 4, 1, s_dump);
    for (int a = 0; a < nctu; a++) {
      int hdr[4]; hdr[0] = (int)TDecCu::bits[a];
      int n[3]; float* src[3] = { TDecCu::CUPU[a], TDecCu::Pred[a], TDecCu::MV[a] };
      for (int s = 0; s < 3; s++) { n[s] = 0; while (src[s][n[s]] != 999) n[s]++; hdr[1+s] = n[s]; }
      fwrite(hdr, 4, 4, s_dump);
      for (int s = 0; s < 3; s++) for (int i = 0; i < n[s]; i++) { short v = (short)src[s][i]; fwrite(&v, 2, 1, s_dump); }
    }
    fflush(s_dump);
    /* packed wire format on an inherited file descriptor (HM_PACKED_FD): the decoder as a producer PROCESS for a GPU-owning
       consumer (compressd_domain_saliencyprediction_b200/feeder.py).  Record: int32 poc, nctu, ntok; int32 hdr[nctu][4] =
       {token offset, n_cupu, n_pred, n_mv}; uint16 bytes[nctu]; int16 tok[ntok] (per CTU: cupu | pred | mv), as include/cds.h */
    static FILE* s_pk = NULL; static int s_pk_init = 0;
    if (!s_pk_init) { s_pk_init = 1; const char* fd = getenv("HM_PACKED_FD"); if (fd) s_pk = fdopen(atoi(fd), "wb"); }
    if (s_pk) {
      static int* ph = NULL; static unsigned short* pb = NULL; static short* pt = NULL; static int cap = 0;
      if (!ph) { ph = (int*)malloc(sizeof(int) * 4 * LCUNUM); pb = (unsigned short*)malloc(2 * LCUNUM); }
      int ntok = 0;
      for (int a = 0; a < nctu; a++) {
        float* src[3] = { TDecCu::CUPU[a], TDecCu::Pred[a], TDecCu::MV[a] };
        ph[4 * a] = ntok;
        for (int s = 0; s < 3; s++) { int n = 0; while (src[s][n] != 999) n++; ph[4 * a + 1 + s] = n; ntok += n; }
        int by = (int)TDecCu::bits[a]; pb[a] = (unsigned short)(by < 0 ? 0 : by > 65535 ? 65535 : by);
      }
      if (ntok > cap) { cap = ntok + 4096; free(pt); pt = (short*)malloc(2 * (size_t)cap); }
      int o = 0;
      for (int a = 0; a < nctu; a++) {
        float* src[3] = { TDecCu::CUPU[a], TDecCu::Pred[a], TDecCu::MV[a] };
        for (int s = 0; s < 3; s++) for (int i = 0; i < ph[4 * a + 1 + s]; i++) pt[o++] = (short)src[s][i];
      }
      int head[3] = { poc, nctu, ntok };
      fwrite(head, 4, 3, s_pk); fwrite(ph, 16, nctu, s_pk); fwrite(pb, 2, nctu, s_pk); fwrite(pt, 2, ntok, s_pk);
      fflush(s_pk);
    }
    static FILE* s_sem = NULL; static int s_sem_init = 0;
    if (!s_sem_init) { s_sem_init = 1; const char* pth = getenv("HM_SEMANTIC_DUMP"); if (pth) s_sem = fopen(pth, "wb"); }
    if (s_sem) {      /* int32 poc, int32 nctu, then nctu x (256 depth, 256 pred mode, 256 x 2 int16 mv) in z-scan order */
      fwrite(&poc, 4, 1, s_sem); fwrite(&nctu, 4, 1, s_sem);
      for (int a = 0; a < nctu; a++) { fwrite(g_cdsSemDepth[a], 1, 256, s_sem); fwrite(g_cdsSemPred[a], 1, 256, s_sem); fwrite(g_cdsSemMv[a], 2, 512, s_sem); }
      fflush(s_sem);
    }
  }
}
'''

# SURVEY 8(f).2 "semantic" capture: the CTU's 256 z-scan partitions as TComDataCU holds them (fixture generation for cds_getframe_semantic)
SEMANTIC_GLOBALS = r'''
unsigned char g_cdsSemDepth[LCUNUM][256]; unsigned char g_cdsSemPred[LCUNUM][256]; short g_cdsSemMv[LCUNUM][256][2];
'''
SEMANTIC_CAPTURE = r'''
  { const UInt a_ = pcCU->getAddr(); const Int np_ = (Int)pcCU->getTotalNumPart();
    for (Int i_ = 0; i_ < np_ && i_ < 256; i_++) {
      const Int dir_ = pcCU->getInterDir(i_);
      TComMv mv_;
      if (pcCU->getPredictionMode(i_) == MODE_INTER) { if (dir_ & 1) mv_ = pcCU->getCUMvField(REF_PIC_LIST_0)->getMv(i_); else if (dir_ & 2) mv_ = pcCU->getCUMvField(REF_PIC_LIST_1)->getMv(i_); }
      g_cdsSemDepth[a_][i_] = pcCU->getDepth(i_); g_cdsSemPred[a_][i_] = (unsigned char)pcCU->getPredictionMode(i_);
      g_cdsSemMv[a_][i_][0] = (short)mv_.getHor(); g_cdsSemMv[a_][i_][1] = (short)mv_.getVer(); } }
'''

STATICS = r'''
#include <stdio.h>
#include <stdlib.h>
float TDecCu::bits[LCUNUM] = {0};
float TDecCu::MV[LCUNUM][1000] = {0};
float TDecCu::CUPU[LCUNUM][1000] = {0};
float TDecCu::Pred[LCUNUM][1000] = {0};
extern Bool g_md5_mismatch;
extern unsigned char g_cdsSemDepth[LCUNUM][256]; extern unsigned char g_cdsSemPred[LCUNUM][256]; extern short g_cdsSemMv[LCUNUM][256][2];
static Void calcAndPrintHashStatus(TComPicYuv& pic, const SEIDecodedPictureHash* pictureHashSEI);
'''

# --mode cds: the same decoder with the saliency block replaced by this repo's binding (include/cds_hm_hook.h):
# the hooks hand packed syntax to libcds_b200.so (dlopen'ed from $CDS_LIB) and write the maps to $CDS_OUT.
CDS_BLOCK = r'''
  /* ---- B200 saliency path (replaces the reference's OpenCV saliency block) ---- */
  {
    CdsHmHook& s_cds_hook = CdsHmHook::get();
    if (getenv("CDS_HOOK_LEGACY")) {      /* the older hand-off through the float statics (kept to compare the two) */
      TComSPS* sps = pcSlice->getSPS();
      Window& cw = sps->getConformanceWindow();
      const int cw_w = (int)sps->getPicWidthInLumaSamples() - cw.getWindowLeftOffset() - cw.getWindowRightOffset();
      const int cw_h = (int)sps->getPicHeightInLumaSamples() - cw.getWindowTopOffset() - cw.getWindowBottomOffset();
      s_cds_hook.onSlice(pcSlice->getPOC(), cw_w, cw_h, (int)rpcPic->getNumCUsInFrame(), TDecCu::bits, &TDecCu::MV[0][0],
                         &TDecCu::Pred[0][0], &TDecCu::CUPU[0][0]);
    } else {
      s_cds_hook.onSliceDone(pcSlice->getPOC());
    }
  }
}
'''

# --mode cds, hook 2: TDecCu::decompressCU hands the CTU's int16 token streams to the binding (no float statics, no 999 scan).
# The reference's MVoutput writes `TDecCu::X[pcCU->getAddr()][i] = v; i++;` (TDecCu.cpp:165-289); in the scratch copy the same
# statements also append to three per-CTU int16 scratch arrays, which decompressCU passes on.
CDS_CTU_GLOBALS = r'''
#include "cds_hm_hook.h"
static short g_cdsCupu[128], g_cdsPred[320], g_cdsMv[1400];
'''
CDS_CTU_CALL = r'''
  if (!getenv("CDS_HOOK_LEGACY")) {
    TComSPS* sps_ = pcCU->getSlice()->getSPS();
    Window& cw_ = sps_->getConformanceWindow();
    const int w_ = (int)sps_->getPicWidthInLumaSamples() - cw_.getWindowLeftOffset() - cw_.getWindowRightOffset();
    const int h_ = (int)sps_->getPicHeightInLumaSamples() - cw_.getWindowTopOffset() - cw_.getWindowBottomOffset();
    CdsHmHook::get().onCtu(w_, h_, (int)pcCU->getPic()->getNumCUsInFrame(), (int)pcCU->getAddr(), (int)pcCU->getTotalBits(),
                           g_cdsCupu, q, g_cdsPred, f, g_cdsMv, p);
  }
'''


def main():
    mode = "cds" if "--mode" in sys.argv and sys.argv[sys.argv.index("--mode") + 1] == "cds" else "dump"
    global WORK
    if mode == "cds":
        WORK = WORK + "_cds"
    if not os.path.isdir(REF):
        print("reference tree absent; nothing to build"); return 0
    if os.path.isdir(WORK): shutil.rmtree(WORK)
    os.makedirs(WORK)
    for d in ("source", "build"):
        shutil.copytree(os.path.join(REF, d), os.path.join(WORK, d))
    subprocess.check_call(["chmod", "-R", "u+w", WORK])
    # (i) -Werror
    mb = os.path.join(WORK, "build/linux/common/makefile.base")
    s = open(mb).read().replace("-Werror", ""); open(mb, "w").write(s)
    # (ii) system("pause")
    for f in ("source/App/TAppDecoder/TAppDecTop.cpp", "source/App/TAppDecoder/decmain.cpp"):
        p = os.path.join(WORK, f); s = open(p, encoding="latin-1").read()
        s = s.replace('system("pause");', ''); open(p, "w", encoding="latin-1").write(s)
    # geometry macro: one binary for every resolution (only sizes static arrays)
    p = os.path.join(WORK, "source/Lib/TLibDecoder/TDecCu.h")
    s = open(p, encoding="latin-1").read()
    s = re.sub(r"#define LCUNUM\s+\d+", "#define LCUNUM 2100", s); open(p, "w", encoding="latin-1").write(s)
    # parse-only producer: with CDS_PARSE_ONLY=1 in the environment the decoder still parses every CTU (CABAC, motion
    # derivation, the hooks) but skips pixel reconstruction, deblocking and SAO - the saliency path needs syntax only
    s = open(p, encoding="latin-1").read()
    s = s.replace("#define FRAMERATE", "#include <stdlib.h>\nstatic inline bool cdsParseOnly() { static int v = -1; if (v < 0) { const char* e = getenv(\"CDS_PARSE_ONLY\"); v = (e && atoi(e)) ? 1 : 0; } return v == 1; }\n#define FRAMERATE", 1)
    assert "cdsParseOnly" in s
    open(p, "w", encoding="latin-1").write(s)
    pc = os.path.join(WORK, "source/Lib/TLibDecoder/TDecCu.cpp")
    s = open(pc, encoding="latin-1").read()
    assert "  xDecompressCU( pcCU, 0,  0 );" in s
    s = s.replace("  xDecompressCU( pcCU, 0,  0 );", "  if (!cdsParseOnly()) xDecompressCU( pcCU, 0,  0 );\n" + SEMANTIC_CAPTURE, 1)
    s = s.replace("Void TDecCu::decompressCU( TComDataCU* pcCU )", SEMANTIC_GLOBALS + "Void TDecCu::decompressCU( TComDataCU* pcCU )", 1)
    if mode == "cds":
        # every capture statement of MVoutput also appends the int16 token; decompressCU then hands the CTU over
        n0 = s.count("TDecCu::MV[pcCU->getAddr()][p] = ")
        s = re.sub(r"TDecCu::MV\[pcCU->getAddr\(\)\]\[p\] = ([^;]+); p\+\+;", r"{ g_cdsMv[p] = (short)(\1); TDecCu::MV[pcCU->getAddr()][p] = \1; p++; }", s)
        s = re.sub(r"TDecCu::CUPU\[pcCU->getAddr\(\)\]\[q\] = ([^;]+);\s*q\+\+;", r"{ g_cdsCupu[q] = (short)(\1); TDecCu::CUPU[pcCU->getAddr()][q] = \1; q++; }", s)
        s = re.sub(r"TDecCu::Pred\[pcCU->getAddr\(\)\]\[f\] = ([^;]+); f\+\+;", r"{ g_cdsPred[f] = (short)(\1); TDecCu::Pred[pcCU->getAddr()][f] = \1; f++; }", s)
        assert s.count("g_cdsMv[p] = ") >= 13 and s.count("g_cdsCupu[q] = ") == 2 and s.count("g_cdsPred[f] = ") == 1, (n0, s.count("g_cdsMv[p] = "), s.count("g_cdsCupu[q] = "), s.count("g_cdsPred[f] = "))
        s = s.replace("  TDecCu::Pred[pcCU->getAddr()][f] = 999;", "  TDecCu::Pred[pcCU->getAddr()][f] = 999;\n" + CDS_CTU_CALL, 1)
        assert "onCtu(" in s
        s = s.replace("Void TDecCu::decompressCU( TComDataCU* pcCU )", CDS_CTU_GLOBALS + "Void TDecCu::decompressCU( TComDataCU* pcCU )", 1)
    open(pc, "w", encoding="latin-1").write(s)
    # (iii) TDecGop.cpp: keep 1-46, 119-224, 337-end (1-based)
    p = os.path.join(WORK, "source/Lib/TLibDecoder/TDecGop.cpp")
    L = open(p, encoding="latin-1").read().replace("\r", "").lstrip("﻿").lstrip("\xef\xbb\xbf").split("\n")
    head = [l for l in L[:46] if "Math.h" not in l]
    mid = L[118:224]
    tail = L[336:472]
    # drop prototypes of the removed OpenCV helpers from the middle block
    mid = [l for l in mid if "Mat" not in l and "vote_alg" not in l and "Umean" not in l
           and "MergeSort" not in l and "Merge(" not in l and "MeanArry" not in l
           and "int max(int" not in l and "g_md5_mismatch" not in l
           and not l.startswith("static Void calcAndPrintHashStatus")]
    block, extra = DUMP, ""
    if mode == "cds":
        block, extra = CDS_BLOCK, '#include "cds_hm_hook.h"\n'
        for hname in ("cds.h", "cds_hm_hook.h"):          # this repo's own headers into the scratch tree
            shutil.copy(os.path.join(HERE, "..", "..", "include", hname), os.path.join(WORK, "source/Lib/TLibDecoder", hname))
    tail_s = "\n".join(tail)
    assert "  m_pcLoopFilter->loopFilterPic( rpcPic );" in tail_s and "  if( pcSlice->getSPS()->getUseSAO() )" in tail_s
    tail_s = tail_s.replace("  m_pcLoopFilter->loopFilterPic( rpcPic );", "  if (!cdsParseOnly()) m_pcLoopFilter->loopFilterPic( rpcPic );", 1)
    tail_s = tail_s.replace("  if( pcSlice->getSPS()->getUseSAO() )", "  if( !cdsParseOnly() && pcSlice->getSPS()->getUseSAO() )", 1)
    out = "\n".join(head) + "\n#include <math.h>\n" + extra + STATICS + "\n".join(mid) + block + tail_s + "\n"
    open(p, "w", encoding="latin-1").write(out)
    env = dict(os.environ)
    for lib in ("TLibVideoIO", "TLibCommon", "TLibDecoder", "TAppCommon"):
        subprocess.check_call(["make", "-j8", "-C", os.path.join(WORK, "build/linux/lib", lib), "release"], env=env)
    subprocess.check_call(["make", "-C", os.path.join(WORK, "build/linux/app/TAppDecoder"), "release"], env=env)
    os.makedirs(OUT, exist_ok=True)
    name = "hm_cds_decoder" if mode == "cds" else "hm_syntax_dump"
    shutil.copy(os.path.join(WORK, "bin/TAppDecoderStatic"), os.path.join(OUT, name))
    print("built", os.path.join(OUT, name))
    return 0

if __name__ == "__main__":
    sys.exit(main())
