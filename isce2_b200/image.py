"""Raw raster + ISCE XML + VRT reader/writer: the part of ``isceobj.Image`` the zero-Doppler components touch.

On-disk contract: headerless native-endian rasters, BIL for the two-band angle layers
(components/iscesys/ImageApi/InterleavedAccessor/src/BILAccessor.cpp:11-37), an ``<imageFile>`` XML property tree
(components/isceobj/Image/Image.py:66-197,374-397,764-819) and a ``VRTRawRasterBand`` VRT (Image.py:461-580).
A real ``isceobj`` image object can be passed wherever one of these is expected: only its metadata getters
(filename/width/length/bands/dataType/scheme and, for DEMs, first/delta latitude/longitude) are used.
"""
from __future__ import annotations

import contextlib
import os
import xml.etree.ElementTree as ET

import numpy as np

# Image.py:62-63
TO_NUMPY = {"BYTE": np.int8, "SHORT": np.int16, "INT": np.int32, "LONG": np.int64, "FLOAT": np.float32,
            "DOUBLE": np.float64, "CFLOAT": np.complex64, "CDOUBLE": np.complex128}
VRT_TYPE = {"BYTE": "Byte", "SHORT": "Int16", "INT": "Int32", "LONG": "Int64", "FLOAT": "Float32", "DOUBLE": "Float64",
            "CFLOAT": "CFloat32", "CDOUBLE": "CFloat64"}
SIZE = {"BYTE": 1, "SHORT": 2, "INT": 4, "LONG": 8, "FLOAT": 4, "DOUBLE": 8, "CFLOAT": 8, "CDOUBLE": 16}
ISCE_VERSION = "b200-zerodop-geometry (ISCE2-compatible raster metadata)"


class Coordinate:
    """Image.py:764-819 ImageCoordinate"""

    def __init__(self, start=0.0, delta=1.0, size=None):
        self.coordStart = start
        self.coordDelta = delta
        self.coordSize = size

    @property
    def coordEnd(self):
        if self.coordSize is None:
            return None
        return self.coordStart + self.coordSize * self.coordDelta


class Image:
    family = "image"

    def __init__(self, family=None, name=""):
        self.family = family or self.__class__.family
        self.name = name
        self.filename = ""
        self.width = None
        self.length = None
        self.bands = 1
        self.dataType = "FLOAT"
        self.scheme = "BIP"
        self.accessMode = "read"
        self.byteOrder = "l"
        self.imageType = ""
        self.description = ""
        self.xmin = None
        self.xmax = None
        self.coord1 = Coordinate()
        self.coord2 = Coordinate()
        self.reference = None
        self.caster = None
        self._mmap = None

    # ---- setters / getters used by callers (Image.py:633-760) ----
    def setFilename(self, v): self.filename = v
    def getFilename(self): return self.filename
    def setWidth(self, v): self.width = int(v)
    def getWidth(self): return self.width
    def setLength(self, v): self.length = int(v)
    def getLength(self): return self.length
    def setBands(self, v): self.bands = int(v)
    def getBands(self): return self.bands
    def setDataType(self, v): self.dataType = str(v).upper()
    def getDataType(self): return self.dataType
    def setInterleavedScheme(self, v): self.scheme = str(v).upper()
    def getInterleavedScheme(self): return self.scheme
    def setAccessMode(self, v): self.accessMode = str(v).lower()
    def getAccessMode(self): return self.accessMode
    def setImageType(self, v): self.imageType = v
    def getImageType(self): return self.imageType
    def addDescription(self, v): self.description = v
    def setXmin(self, v): self.xmin = v
    def setXmax(self, v): self.xmax = v
    def getImagePointer(self): return 0  # no DataAccessor handles behind this implementation

    def setCaster(self, mode, dataType):
        """Image.py:594-606: records the in-memory type wanted by the consumer (e.g. DEM 'read' as FLOAT)."""
        self.caster = (mode, str(dataType).upper())

    def initImage(self, filename, accessmode, width, dataType=None, bands=None, scheme=None, caster=None):
        """Image.py:298-317"""
        self.filename = filename
        self.accessMode = str(accessmode).lower()
        self.width = int(width)
        if dataType:
            self.dataType = str(dataType).upper()
        if bands:
            self.bands = int(bands)
        if scheme:
            self.scheme = str(scheme).upper()

    # ---- raster access ----
    def shape(self):
        """Image.py:319-335 memMap shapes"""
        if self.bands == 1:
            return (self.length, self.width)
        s = self.scheme.upper()
        if s == "BIL":
            return (self.length, self.bands, self.width)
        if s == "BIP":
            return (self.length, self.width, self.bands)
        return (self.bands, self.length, self.width)

    def _dtype(self):
        """numpy dtype of one sample in the file's byte order (BYTE_ORDER 'l' / 'b', Image.py:62-63; the reference's
        DataAccessor swaps on the fly, here the memmap carries the order and consumers convert on read)."""
        return _file_dtype(self.dataType, self.byteOrder)

    def createImage(self):
        """Open (read) or create (write) the raster as a numpy memmap."""
        dt = self._dtype()
        if self.accessMode.startswith("r"):
            if self.length is None:
                self.length = os.path.getsize(self.filename) // (dt.itemsize * self.width * self.bands)
            self._mmap = np.memmap(self.filename, dtype=dt, mode="r", shape=self.shape())
        else:
            if self.length is None:
                raise ValueError("length must be set before creating a write-mode image")
            d = os.path.dirname(os.path.abspath(self.filename))
            os.makedirs(d, exist_ok=True)
            self._mmap = np.memmap(self.filename, dtype=dt, mode="w+", shape=self.shape())
        self.coord1.coordSize = self.width
        self.coord2.coordSize = self.length
        return self._mmap

    def memMap(self, mode="r", band=None):
        if self._mmap is None:
            self.createImage()
        if band is None:
            return self._mmap
        s = self.scheme.upper()
        if self.bands == 1:
            return self._mmap
        return self._mmap[:, band, :] if s == "BIL" else (self._mmap[:, :, band] if s == "BIP" else self._mmap[band])

    def asarray(self):
        return np.asarray(self.memMap())

    def toNumpyDataType(self):
        """Image.py:255-256"""
        return TO_NUMPY[self.dataType.upper()]

    def clone(self):
        """A metadata copy without the open raster (Image.clone as mroipac/looks/Looks.py:38,45 uses it)."""
        import copy
        mm, self._mmap = self._mmap, None
        try:
            c = copy.deepcopy(self)
        finally:
            self._mmap = mm
        return c

    def finalizeImage(self):
        # dropping the map is what closing the reference's ofstream is: dirty pages reach the file through the page cache;
        # no msync (it would force synchronous write-back of the whole raster on a disk file system)
        self._mmap = None

    # ---- metadata ----
    def _props(self):
        p = [("ISCE_VERSION", ISCE_VERSION, None),
             ("access_mode", self.accessMode, "Image access mode."),
             ("byte_order", self.byteOrder, "Endianness of the image."),
             ("data_type", self.dataType.upper(), "Image data type."),
             ("family", self.family, "Instance family name"),
             ("file_name", self.filename, "Name of the image file."),
             ("image_type", self.imageType, "Image type used for displaying."),
             ("length", self.length, "Image length"),
             ("name", f"{self.family}_name", "Instance name"),
             ("number_bands", self.bands, "Number of image bands."),
             ("scheme", self.scheme.upper(), "Interleaving scheme of the image."),
             ("width", self.width, "Image width")]
        if self.description:
            p.append(("description", self.description, "Image description"))
        if self.reference is not None:
            p.append(("reference", self.reference, "Geodetic datum"))
        if self.xmin is not None:
            p.append(("xmin", self.xmin, "Minimum range value"))
        if self.xmax is not None:
            p.append(("xmax", self.xmax, "Maximum range value"))
        return sorted(p, key=lambda t: t[0].lower())

    def renderHdr(self, outfile=None):
        """Image.py:374-397: <file>.xml (+ <file>.vrt)"""
        root = ET.Element("imageFile")

        def prop(parent, name, value, doc=None):
            e = ET.SubElement(parent, "property", name=name)
            ET.SubElement(e, "value").text = str(value)
            if doc:
                ET.SubElement(e, "doc").text = doc

        def coord(name, c, doc):
            e = ET.SubElement(root, "component", name=name)
            ET.SubElement(e, "factorymodule").text = "isceobj.Image"
            ET.SubElement(e, "factoryname").text = "createCoordinate"
            ET.SubElement(e, "doc").text = doc
            prop(e, "delta", c.coordDelta, "Coordinate quantization.")
            prop(e, "endingvalue", c.coordEnd, "Ending value of the coordinate.")
            prop(e, "family", "imagecoordinate", "Instance family name")
            prop(e, "name", "imagecoordinate_name", "Instance name")
            prop(e, "size", c.coordSize, "Coordinate size.")
            prop(e, "startingvalue", c.coordStart, "Starting value of the coordinate.")

        self.coord1.coordSize = self.width
        self.coord2.coordSize = self.length
        items = [(n.lower(), ("p", n, v, d)) for n, v, d in self._props()]
        items.append(("coordinate1", ("c", "coordinate1", self.coord1, "First coordinate of a 2D image (width).")))
        items.append(("coordinate2", ("c", "coordinate2", self.coord2, "Second coordinate of a 2D image (length).")))
        for _, it in sorted(items, key=lambda t: t[0]):
            if it[0] == "p":
                prop(root, it[1], it[2], it[3])
            else:
                coord(it[1], it[2], it[3])
        _indent(root)
        ET.ElementTree(root).write((outfile or self.filename) + ("" if outfile else ".xml"), encoding="unicode")
        self.renderVRT()

    def renderVRT(self, outfile=None):
        """Image.py:461-580"""
        root = ET.Element("VRTDataset", rasterXSize=str(self.width), rasterYSize=str(self.length))
        trivial = (self.coord1.coordStart == 0.0 and self.coord2.coordStart == 0.0 and self.coord1.coordDelta == 1.0
                   and self.coord2.coordDelta == 1.0)
        if not trivial:
            ET.SubElement(root, "SRS").text = "EPSG:4326"
            ET.SubElement(root, "GeoTransform").text = "{0}, {1}, 0.0, {2}, 0.0, {3}".format(
                self.coord1.coordStart, self.coord1.coordDelta, self.coord2.coordStart, self.coord2.coordDelta)
        nbytes = SIZE[self.dataType.upper()]
        for band in range(self.bands):
            b = ET.SubElement(root, "VRTRasterBand", dataType=VRT_TYPE[self.dataType.upper()], band=str(band + 1),
                              subClass="VRTRawRasterBand")
            ET.SubElement(b, "SourceFilename", relativeToVRT="1").text = os.path.basename(self.filename)
            ET.SubElement(b, "ByteOrder").text = "LSB" if self.byteOrder.lower() == "l" else "MSB"
            s = self.scheme.upper()
            if s == "BIL":
                off, pix, lin = band * self.width * nbytes, nbytes, self.bands * self.width * nbytes
            elif s == "BIP":
                off, pix, lin = band * nbytes, self.bands * nbytes, self.bands * self.width * nbytes
            else:
                off, pix, lin = band * self.width * self.length * nbytes, nbytes, self.width * nbytes
            ET.SubElement(b, "ImageOffset").text = str(off)
            ET.SubElement(b, "PixelOffset").text = str(pix)
            ET.SubElement(b, "LineOffset").text = str(lin)
        _indent(root)
        ET.ElementTree(root).write(outfile or (self.filename + ".vrt"), encoding="unicode")

    def load(self, xmlfile):
        """Image.load: read an ISCE <imageFile> XML (old upper-case and current lower-case property names)."""
        root = ET.parse(xmlfile).getroot()
        props = {}
        for e in root.findall("property"):
            v = e.find("value")
            props[e.get("name").lower()] = v.text.strip() if v is not None and v.text else ""
        for e in root.findall("component"):
            cname = e.get("name").lower()
            c = self.coord1 if cname == "coordinate1" else self.coord2 if cname == "coordinate2" else None
            if c is None:
                continue
            for pe in e.findall("property"):
                v = pe.find("value")
                txt = v.text.strip() if v is not None and v.text else ""
                key = pe.get("name").lower()
                try:
                    if key == "startingvalue":
                        c.coordStart = float(txt)
                    elif key == "delta":
                        c.coordDelta = float(txt)
                    elif key == "size":
                        c.coordSize = int(float(txt))
                except ValueError:
                    pass
        self.width = int(props.get("width", self.width or 0))
        self.length = int(props.get("length", self.length or 0))
        self.bands = int(props.get("number_bands", self.bands))
        self.dataType = props.get("data_type", self.dataType).upper()
        self.scheme = props.get("scheme", self.scheme).upper()
        self.byteOrder = props.get("byte_order", self.byteOrder)
        self.imageType = props.get("image_type", self.imageType)
        # an image described by an existing header is opened for reading unless the caller says otherwise
        # (the XML keeps the mode it was rendered with, usually 'write')
        self.accessMode = "read"
        self.description = props.get("description", "")
        if "reference" in props:
            self.reference = props["reference"]
        fn = props.get("file_name", "")
        base = xmlfile[:-4] if xmlfile.endswith(".xml") else xmlfile
        # ISCE stores the name used at creation time; prefer the raster sitting next to the XML
        self.filename = base if os.path.exists(base) else fn
        return self


class DemImage(Image):
    """components/isceobj/Image/DemImage.py:60-79 + Image.py:711-733"""
    family = "demimage"

    def __init__(self, name=""):
        super().__init__(name=name)
        self.dataType = "SHORT"
        self.imageType = "dem"
        self.reference = "EGM96"

    def getFirstLongitude(self): return self.coord1.coordStart
    def getDeltaLongitude(self): return self.coord1.coordDelta
    def getFirstLatitude(self): return self.coord2.coordStart
    def getDeltaLatitude(self): return self.coord2.coordDelta
    def setFirstLongitude(self, v): self.coord1.coordStart = float(v)
    def setDeltaLongitude(self, v): self.coord1.coordDelta = float(v)
    def setFirstLatitude(self, v): self.coord2.coordStart = float(v)
    def setDeltaLatitude(self, v): self.coord2.coordDelta = float(v)
    firstLongitude = property(getFirstLongitude, setFirstLongitude)
    deltaLongitude = property(getDeltaLongitude, setDeltaLongitude)
    firstLatitude = property(getFirstLatitude, setFirstLatitude)
    deltaLatitude = property(getDeltaLatitude, setDeltaLatitude)


def createImage(name=""):
    return Image(name=name)


def createDemImage(name=""):
    return DemImage(name=name)


def _indent(elem, level=0):
    pad = "\n" + "    " * level
    if len(elem):
        if not elem.text or not elem.text.strip():
            elem.text = pad + "    "
        for child in elem:
            _indent(child, level + 1)
        if not child.tail or not child.tail.strip():
            child.tail = pad
    if level and (not elem.tail or not elem.tail.strip()):
        elem.tail = pad


def _file_dtype(data_type, byte_order="l"):
    dt = np.dtype(TO_NUMPY[str(data_type).upper()])
    bo = str(byte_order or "l").lower()[:1]
    if bo not in ("l", "b"):
        raise ValueError(f"unknown BYTE_ORDER {byte_order!r} (expected 'l' or 'b')")
    return dt.newbyteorder(">" if bo == "b" else "<") if dt.itemsize > 1 else dt


def _foreign_meta(img, need_length=True):
    """(filename, dtype, width, length, bands, scheme) of an image object that is not ours (an isceobj Image: duck-typed)."""
    fn = img.getFilename() if hasattr(img, "getFilename") else img.filename
    dt = _file_dtype(img.dataType, getattr(img, "byteOrder", "l"))
    width, bands = int(img.width), int(getattr(img, "bands", 1) or 1)
    length = getattr(img, "length", None)
    if not length:
        length = os.path.getsize(fn) // (dt.itemsize * width * bands) if need_length else 0
    return fn, dt, width, int(length), bands, str(getattr(img, "scheme", "BIL") or "BIL").upper()


def _shape(length, width, bands, scheme):
    if bands == 1:
        return (length, width)
    return {"BIL": (length, bands, width), "BIP": (length, width, bands)}.get(scheme, (bands, length, width))


def read_raster(img, as_dtype=None):
    """Return the raster behind `img` (ours or an isceobj image: duck-typed metadata) as a numpy array in native byte
    order, in the image's own interleaving (Image.py:319-335 shapes)."""
    if isinstance(img, Image):
        arr = img.memMap()
    else:
        fn, dt, width, length, bands, scheme = _foreign_meta(img)
        arr = np.memmap(fn, dtype=dt, mode="r", shape=_shape(length, width, bands, scheme))
    if not arr.dtype.isnative:
        arr = np.asarray(arr).astype(arr.dtype.newbyteorder("="))
    if as_dtype is not None and arr.dtype != np.dtype(as_dtype):
        arr = np.asarray(arr).astype(as_dtype)
    return arr


def output_memmap(img, length, width, bands=1):
    """Writable [length][(bands)][width] memmap of an OUTPUT image the caller handed to a Component ('Must either pass
    the latImage in the call or set self.latFilename', Topozero.py:274-302).  Ours: its own memmap.  A foreign (isceobj)
    image: its memMap() defaults to read-only and shapes single-band rasters (length, 1, width), so the file it names is
    mapped here instead, created at the right size if need be."""
    if isinstance(img, Image):
        return img.memMap()
    fn, dt, w, _, b, scheme = _foreign_meta(img, need_length=False)
    if w != width or b != bands or (bands > 1 and scheme != "BIL"):
        raise ValueError(f"output image {fn}: {w} samples x {b} bands ({scheme}) where {width} x {bands} (BIL) are written")
    if not dt.isnative:
        raise ValueError(f"output image {fn}: big-endian output rasters are not supported")
    os.makedirs(os.path.dirname(os.path.abspath(fn)), exist_ok=True)
    need = length * width * bands * dt.itemsize
    mode = "r+" if os.path.exists(fn) and os.path.getsize(fn) == need else "w+"
    return np.memmap(fn, dtype=dt, mode=mode, shape=_shape(length, width, bands, "BIL"))


def _file_range(a, writable):
    """(address, nbytes, filename, file offset) when the array `a` is -- or is a contiguous view of -- a numpy.memmap opened
    for writing (writable) or at all, else None."""
    if not isinstance(a, np.ndarray) or a.size == 0 or not a.flags["C_CONTIGUOUS"]:
        return None
    mm = a
    while mm is not None and not isinstance(mm, np.memmap):
        mm = getattr(mm, "base", None)
    # a view of a memmap is itself a memmap whose attributes describe the parent: the mapped buffer proper is the one
    # whose base is the mmap object
    while isinstance(mm, np.memmap) and isinstance(mm.base, np.memmap):
        mm = mm.base
    if not isinstance(mm, np.memmap) or not getattr(mm, "filename", None):
        return None
    if writable and (getattr(mm, "mode", "r") not in ("r+", "w+") or not a.flags["WRITEABLE"]):
        return None
    if getattr(mm, "mode", "r") == "c":  # copy-on-write: the file is not what the mapping shows
        return None
    delta = a.ctypes.data - mm.ctypes.data
    if delta < 0 or delta + a.nbytes > mm.nbytes:
        return None
    return a.ctypes.data, a.nbytes, str(mm.filename), int(mm.offset) + delta


@contextlib.contextmanager
def file_backed(arrays, inputs=()):
    """For the duration of a library call: every writable numpy.memmap among `arrays` (None entries allowed) is declared
    to the library as the file mapping it is (b200_host_file_register), so that results are written into the file with
    pwrite, one 32 MB slot per call, instead of through the mapping -- the same pages, without a page fault per 4 KB of a
    raster that does not exist yet (16.5 GB swath on tmpfs: 0.9 s against 1.8 s, profiles/r02_component_file_sweep.log).
    `inputs`: arrays the call READS that are (views of) memmaps of rasters; they are uploaded with pread.
    B200_FILE_WRITES=0 in the environment leaves everything to the mappings."""
    from . import _capi
    done = []
    try:
        if os.environ.get("B200_FILE_WRITES", "1") != "0":
            for a, writable in [(a, True) for a in arrays] + [(a, False) for a in inputs]:
                r = _file_range(a, writable)
                if r is None or r[0] in done:
                    continue
                addr, nbytes, filename, offset = r
                try:
                    fd = os.open(filename, os.O_RDWR if writable else os.O_RDONLY)
                except OSError:
                    continue
                try:
                    _capi.host_file_register(addr, nbytes, fd, offset)
                    done.append(addr)
                except _capi.B200Error:  # e.g. two images over one file: the mapping still works
                    pass
                finally:
                    os.close(fd)  # the library holds its own duplicate
        yield
    finally:
        for addr in done:
            _capi.host_file_unregister(addr)


def read_view(path):
    """Open the raster named `path` (no extension; ``path.xml`` / ``path.vrt`` beside it).  When the raster file itself
    does not exist and ``path.vrt`` is a window into a parent raster (the per-burst views written by
    TopsProc/runTopo.py:362-423 ``buildVRT``: one ``SimpleSource`` + ``SrcRect`` per band), the window of the parent
    is returned without copying; the reference reads such views through GDAL (DataAccessorPy.py:125-171)."""
    if path.endswith(".xml") or path.endswith(".vrt"):
        path = path[:-4]
    if os.path.exists(path) and os.path.exists(path + ".xml"):
        img = Image()
        img.load(path + ".xml")
        img.filename = path
        return img.memMap()
    if not os.path.exists(path + ".vrt"):
        raise FileNotFoundError(path)
    root = ET.parse(path + ".vrt").getroot()
    views = []
    for b in root.findall("VRTRasterBand"):
        src = b.find("SimpleSource")
        if src is None:  # a raw VRT without its XML: describe the file from the band record
            raise FileNotFoundError(path + ".xml")
        fn = src.find("SourceFilename")
        parent = fn.text.strip()
        if fn.get("relativeToVRT", "0") == "1":
            parent = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(path)), parent))
        rect = src.find("SrcRect")
        x0, y0 = int(float(rect.get("xOff"))), int(float(rect.get("yOff")))
        nx, ny = int(float(rect.get("xSize"))), int(float(rect.get("ySize")))
        band = int(src.find("SourceBand").text) - 1
        arr = read_view(parent)
        if arr.ndim == 3:
            arr = arr[:, band, :]  # BIL parents only (the layers topo writes)
        views.append(arr[y0:y0 + ny, x0:x0 + nx])
    if len(views) == 1:
        return views[0]
    return np.stack(views, axis=1)
