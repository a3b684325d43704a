#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — builds oracle/_ref/libplf_ref.so from the REFERENCE'S OWN SOURCES where they lie under
/root/reference (nothing of them is copied into this repository; oracle/_ref/ is git-ignored).

What is compiled, unmodified:
  whole files  src/ORBextractor.cc  src/LineExtractor.cc  src/LineMatcher.cpp  src/gridStructure.cpp
               src/LineIterator.cpp  src/Config.cpp  Thirdparty/line_descriptor/src/LSDDetector_custom.cpp
               Thirdparty/DBoW2/DBoW2/{FORB,BowVector,FeatureVector,ScoringObject}.cpp  Thirdparty/DBoW2/DUtils/*.cpp
               (+ the header-only TemplatedVocabulary.h, instantiated for FORB in refshim/ref_capi.cpp)
  line ranges  (written at build time to oracle/_ref/gen/, never committed)
               src/Frame.cc:976-1307          ComputeStereoMatches, ComputeStereoMatches_Lines,
                                              lineSegmentOverlapStereo, filterLineSegmentDisparity
               src/ORBmatcher.cc:36-42,2495-2511   TH_HIGH / TH_LOW, DescriptorDistance
               src/ORBmatcher.cc:269-471,2449-2490 SearchByBoW(KeyFrame*, Frame&, ...), ComputeThreeMaxima
               src/ORBmatcher.cc:44-222,473-586,2179-2447  the four SearchByProjection overloads + RadiusByViewingCos
               src/Frame.cc:451-482,774-855   AssignFeaturesToGrid, GetFeaturesInArea, PosInGrid
               src/KeyFrame.cc:881-930        KeyFrame::GetFeaturesInArea, IsInImage
               src/Tracking.cc:3055-3099,3879-3919 match() + the orientation / position gates of TrackWithMotionModel
                                              and SearchLocalLines (loop bodies wrapped in two functions)
               Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:42-687,1026-1372
                                              LBD: parameters, computeSobel, binaryConversion, computeImpl, computeLBD
                                              (the rest of that file is the EDLine detector, dead on this path)
against oracle/refshim/: a stand-in for the OpenCV / Eigen headers (neither is installed here) whose arithmetic
primitives forward to the cv2-pinned restatements of oracle/cpp (prims.cpp, orb.cpp fast_window, lsd.cpp lsd_detect).
The reference's own build system (cmake + OpenCV 3 + Eigen + Pangolin + DBoW2 + g2o) is not run.

Flags follow the oracle: -O2 -ffp-contract=off (float expressions evaluated as written).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLF_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
LIB = os.path.join(OUT, "libplf_ref.so")          # the pin: -O2 -ffp-contract=off, float expressions as written
LIB_O3 = os.path.join(OUT, "libplf_ref_o3.so")    # the timed build: -O3 -march=x86-64-v3 (the reference's CMake uses -O3 -march=native)
SHIM = os.path.join(HERE, "refshim")
LD = os.path.join(REF, "Thirdparty", "line_descriptor")

WHOLE = ["src/ORBextractor.cc", "src/LineExtractor.cc", "src/LineMatcher.cpp", "src/gridStructure.cpp",
         "src/LineIterator.cpp", "src/Config.cpp", "Thirdparty/line_descriptor/src/LSDDetector_custom.cpp",
         # DBoW2 as vendored by the reference (bag-of-words transform of Frame::ComputeBoW): the whole library
         "Thirdparty/DBoW2/DBoW2/FORB.cpp", "Thirdparty/DBoW2/DBoW2/BowVector.cpp", "Thirdparty/DBoW2/DBoW2/FeatureVector.cpp",
         "Thirdparty/DBoW2/DBoW2/ScoringObject.cpp", "Thirdparty/DBoW2/DUtils/Random.cpp", "Thirdparty/DBoW2/DUtils/Timestamp.cpp"]
ORACLE_PRIMS = ["cpp/prims.cpp", "cpp/orb.cpp", "cpp/lsd.cpp"]   # pinned OpenCV primitives (orb/lsd also carry oracle logic, unused here)

# The system compiler when it is there: this image's $CXX is a wrapper that links libstdc++ STATICALLY, and a second copy
# of libstdc++ inside a dlopen'ed library breaks iostream extraction (locale facet ids are GNU-unique symbols shared with
# the process's libstdc++.so.6) - DBoW2's loadFromTextFile parses with operator>>.
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")
BASE = ["-O2", "-std=c++14", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w", "-fvisibility=hidden"]
INC = ["-I", SHIM, "-I", os.path.join(REF, "include"), "-I", os.path.join(LD, "include"), "-I", os.path.join(LD, "src"),
       "-I", os.path.join(REF, "Thirdparty", "DBoW2"), "-I", os.path.join(REF, "Thirdparty", "DBoW2", "DBoW2")]
# src/ files that reach include/Frame.h in the reference see ref_types.h instead (Frame / ORBmatcher / MapLine stand-ins);
# the line_descriptor library is compiled exactly as its own CMakeLists does: its own precomp header, nothing else
# (no using-directives leak in - that decides e.g. which atan2 overload LSDDetector_custom.cpp:297 binds to).
FRAME_SIDE = ["-DFRAME_H", "-DORBMATCHER_H", "-DMAPPOINT_H", "-DMAPLINE_H", "-DKEYFRAME_H",
              "-include", os.path.join(SHIM, "ref_types.h")]


def lines(path, a, b):
    # newline="\n": line numbers as grep / sed / an editor count them (a lone CR inside a line is not a line break)
    with open(os.path.join(REF, path), encoding="utf-8", errors="replace", newline="\n") as f:
        ls = f.readlines()
    return "".join(ls[a - 1:b]).replace("\r", "")


def generate():
    os.makedirs(GEN, exist_ok=True)
    frame = ("// GENERATED at build time from %s/src/Frame.cc:976-1307 and src/ORBmatcher.cc:36-42,2495-2511 - do not commit\n"
             "namespace ORB_SLAM3 {\n" % REF
             + lines("src/ORBmatcher.cc", 36, 42) + lines("src/ORBmatcher.cc", 2495, 2511)
             + lines("src/Frame.cc", 976, 1307)
             + "\n// ORBmatcher::SearchByBoW (src/ORBmatcher.cc:269-471) and ComputeThreeMaxima (:2449-2490)\n"
             + lines("src/ORBmatcher.cc", 269, 471) + lines("src/ORBmatcher.cc", 2449, 2490)
             + "\n// the SearchByProjection overloads: local map (:44-214 with RadiusByViewingCos :216-222), loop closing (:473-586),\n"
             "// frame to frame (:2179-2323), relocalisation (:2325-2447)\n"
             + lines("src/ORBmatcher.cc", 44, 222) + lines("src/ORBmatcher.cc", 473, 586) + lines("src/ORBmatcher.cc", 2179, 2447)
             + "\n// Frame::AssignFeaturesToGrid (src/Frame.cc:451-482), GetFeaturesInArea (:774-843), PosInGrid (:845-855);\n"
             "// KeyFrame::GetFeaturesInArea, IsInImage (src/KeyFrame.cc:881-930)\n"
             + lines("src/Frame.cc", 451, 482) + lines("src/Frame.cc", 774, 855) + lines("src/KeyFrame.cc", 881, 930)
             + "\nfloat Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;\n"
             "float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;\n"
             "// Tracking::TrackWithMotionModel, line half: src/Tracking.cc:3055-3099\n"
             "int ref_track_gate_f2f(Frame& mCurrentFrame, Frame& mLastFrame, std::vector<int>& matches_out) {\n"
             + lines("src/Tracking.cc", 3055, 3099)
             + "    matches_out = matches_12;\n    return mCurrentFrame.n_inliers_ls;\n}\n"
             "// Tracking::SearchLocalLines, matching half: src/Tracking.cc:3879-3919\n"
             "void ref_track_gate_local(Frame& mCurrentFrame, std::vector<MapLine*>& mvpLocalMapLines_InFrustum, int nToMatch, std::vector<int>& matches_out) {\n"
             + lines("src/Tracking.cc", 3879, 3919)
             + "    matches_out = matches_12;\n}\n"
             + "\n}\n")
    lbd = ("// GENERATED at build time from %s/Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:42-687,1026-1372 - do not commit\n" % REF
           + lines("Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp", 42, 687)
           + lines("Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp", 1026, 1372)
           + "\n// members of the EDLine half of the file, named by the ranges above but never reached on this path\n"
           "int BinaryDescriptor::OctaveKeyLines( cv::Mat&, ScaleLines& ) { throw std::runtime_error(\"EDLine is not built\"); }\n"
           "BinaryDescriptor::EDLineDetector::EDLineDetector() {}\n"
           "BinaryDescriptor::EDLineDetector::~EDLineDetector() {}\n"
           "}\n}\n")
    for name, text in (("frame_ranges.cpp", frame), ("lbd_ranges.cpp", lbd)):
        p = os.path.join(GEN, name)
        if not os.path.exists(p) or open(p).read() != text:
            with open(p, "w") as f:
                f.write(text)
    return [os.path.join(GEN, "frame_ranges.cpp"), os.path.join(GEN, "lbd_ranges.cpp")]


def build(force=False):
    a = build_one(LIB, ["-O2", "-ffp-contract=off", "-fno-fast-math"], "obj", force)
    build_one(LIB_O3, ["-O3", "-march=x86-64-v3"], "obj_o3", force)
    return a


def build_one(LIB, OPT, objname, force=False):
    if not os.path.isdir(os.path.join(REF, "src")):
        if os.path.exists(LIB):
            return LIB            # GPU box: the prebuilt library travelled with the snapshot
        raise RuntimeError("reference tree not found at %s and no prebuilt %s" % (REF, LIB))
    gen = generate()
    srcs = [os.path.join(REF, s) for s in WHOLE] + gen + \
           [os.path.join(SHIM, "cvshim.cpp"), os.path.join(SHIM, "ref_capi.cpp")] + \
           [os.path.join(HERE, s) for s in ORACLE_PRIMS]
    deps = srcs + [os.path.join(dp, f) for dp, _, fs in os.walk(SHIM) for f in fs] + \
           [os.path.join(HERE, "cpp", f) for f in os.listdir(os.path.join(HERE, "cpp")) if f.endswith(".h")] + [__file__]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    objdir = os.path.join(OUT, objname)
    BASE = OPT + ["-std=c++14", "-fPIC", "-w", "-fvisibility=hidden"]
    os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if src.startswith(HERE + os.sep + "cpp"):      # the oracle's own files: no reference headers, C++17
            flags = OPT + ["-std=c++17", "-fPIC", "-w", "-fvisibility=hidden", "-DPLF_ORACLE_BUILD"]
        elif "line_descriptor" in src or "DBoW2" in src or src.endswith("lbd_ranges.cpp") or src.endswith("cvshim.cpp") \
                or os.path.basename(src) in ("gridStructure.cpp", "LineIterator.cpp", "Config.cpp"):
            flags = BASE + INC
        else:
            flags = BASE + INC + FRAME_SIDE
        subprocess.check_call([CXX] + flags + ["-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, srcs))
    subprocess.check_call([CXX, "-shared", "-o", LIB] + objs + ["-Wl,-Bsymbolic", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
