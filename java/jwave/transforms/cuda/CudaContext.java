package jwave.transforms.cuda;

import java.lang.foreign.MemorySegment;

/** Process-wide jwc_ctx (lazily created on the current CUDA device; override with -Djwave.cuda.devices=0,1,...). */
final class CudaContext {
  private static volatile MemorySegment ctx;

  static MemorySegment get() {
    MemorySegment c = ctx;
    if (c == null) {
      synchronized (CudaContext.class) {
        c = ctx;
        if (c == null) {
          String devs = System.getProperty("jwave.cuda.devices");
          int[] d = null;
          if (devs != null && !devs.isBlank()) {
            String[] parts = devs.split(",");
            d = new int[parts.length];
            for (int i = 0; i < parts.length; i++) d[i] = Integer.parseInt(parts[i].trim());
          }
          ctx = c = JwcNative.create(d);
        }
      }
    }
    return c;
  }

  private CudaContext() { }
}
